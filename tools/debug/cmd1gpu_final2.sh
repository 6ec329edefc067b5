mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_pytest_gpu_final.log
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?; tail -3 gpurun_out/r2_bench_n1.err
echo "== ncu c3 lanes"; timeout 600 ncu --set full --clock-control none -k regex:ell_thread -s 4 -c 1 -o gpurun_out/r2_c3_lanes python tools/profile_target.py --config c3 --mode accumulate > gpurun_out/ncu_c3.log 2>&1; echo rc=$?
echo "== skew"; for ml in 200000 50000 20000; do timeout 600 python tools/skew_bench.py --max-len $ml >> gpurun_out/r2_skew_bench.jsonl 2>gpurun_out/skew.err || tail -3 gpurun_out/skew.err; done
echo "== long row floor"; timeout 600 python tools/one_long_row.py > gpurun_out/r2_one_long_row.jsonl 2>gpurun_out/olr.err || tail -3 gpurun_out/olr.err; cat gpurun_out/r2_one_long_row.jsonl
for f in gpurun_out/*.ncu-rep; do b=${f%.ncu-rep}; ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null; ncu -i $f --page details 2>/dev/null | head -c 60000 > ${b}_details.txt; rm -f $f; done
du -sh gpurun_out
