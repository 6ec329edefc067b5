mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_gpu.log 2>&1; echo rc=$?; tail -4 gpurun_out/s3_pytest_gpu.log
echo "== bench N=1 (default command)"; T0=$(date +%s); timeout 900 python bench.py > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench_n1.err; echo rc=$? wall=$(( $(date +%s) - T0 ))s; tail -3 gpurun_out/s3_bench_n1.err
echo "== host csrspmv synthetic"; LC_ALL=C timeout 300 ellspmv_b200/host/bin/csrspmv -v --synthetic=laplace2d:8192,8192 --repeat=3 --warmup=2 -q 2>&1 | tail -3
echo "== odd K"; timeout 600 python tools/odd_k.py > gpurun_out/s3_odd_k.jsonl 2> gpurun_out/s3_odd_k.err; echo rc=$?; cat gpurun_out/s3_odd_k.jsonl; tail -3 gpurun_out/s3_odd_k.err
echo "== ncu c2 csr"; timeout 600 ncu --set full --clock-control none -k regex:ell_thread -s 4 -c 1 -o gpurun_out/r2_c2_csr python tools/profile_target.py --config c2 --path csr > gpurun_out/ncu_c2csr.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_c2csr.log
for f in gpurun_out/*.ncu-rep; do b=${f%.ncu-rep}; ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null; ncu -i $f --page details 2>/dev/null | head -c 60000 > ${b}_details.txt; rm -f $f; done
du -sh gpurun_out
