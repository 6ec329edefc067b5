mkdir -p gpurun_out
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1.err
echo "== bench N=1 200 steps"; timeout 900 python bench.py --steps 200 --warmup 5 --no-other-configs > gpurun_out/r2_bench_n1_200.json 2> gpurun_out/r2_bench_n1_200.err; echo rc=$?
echo "== reference N=1"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err; echo rc=$?
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed
echo "== ncu c2 overwrite"; timeout 600 ncu --set full --clock-control none -k regex:ell_thread -s 4 -c 1 -o gpurun_out/r2_c2_iterate python tools/profile_target.py --config c2 --mode overwrite > gpurun_out/ncu_c2o.log 2>&1; echo rc=$?
echo "== ncu c2 accumulate"; timeout 600 ncu --set full --clock-control none -k regex:ell_thread -s 4 -c 1 -o gpurun_out/r2_c2_accumulate python tools/profile_target.py --config c2 --mode accumulate > gpurun_out/ncu_c2a.log 2>&1; echo rc=$?
echo "== ncu c3"; timeout 600 ncu --set full --clock-control none -k regex:ell_thread -s 4 -c 1 -o gpurun_out/r2_c3 python tools/profile_target.py --config c3 --mode accumulate > gpurun_out/ncu_c3.log 2>&1; echo rc=$?
echo "== ncu c4 ell (metrics)"; timeout 900 ncu --metrics $M --clock-control none -k regex:sg_ -s 27 -c 9 --csv --log-file gpurun_out/r2_c4_staged_launches.csv python tools/profile_target.py --config c4 --launches 2 > gpurun_out/ncu_c4.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_c4.log
echo "== ncu c4 csr (metrics)"; timeout 900 ncu --metrics $M --clock-control none -k regex:sg_ -s 27 -c 9 --csv --log-file gpurun_out/r2_c4_csr_launches.csv python tools/profile_target.py --config c4 --path csr --launches 2 > gpurun_out/ncu_c4csr.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_c4csr.log
echo "== ncu c4 gather kernel full"; timeout 900 ncu --set full --clock-control none -k regex:sg_gather -s 30 -c 1 -o gpurun_out/r2_c4_gather python tools/profile_target.py --config c4 --launches 2 > gpurun_out/ncu_c4g.log 2>&1; echo rc=$?
echo "== host programs"; LC_ALL=C timeout 120 ellspmv_b200/host/bin/ellspmv -v --synthetic=laplace2d:8192,8192 --repeat=5 -q 2>&1 | tail -3
for f in gpurun_out/*.ncu-rep; do b=${f%.ncu-rep}; ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null; ncu -i $f --page details 2>/dev/null | head -c 60000 > ${b}_details.txt; rm -f $f; done
du -sh gpurun_out
