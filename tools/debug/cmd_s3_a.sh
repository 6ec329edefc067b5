mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_gpu.log 2>&1; echo rc=$?; tail -15 gpurun_out/s3_pytest_gpu.log
echo "== bench N=1"; /usr/bin/time -v -o gpurun_out/s3_bench_time.txt timeout 900 python bench.py > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench_n1.err; echo rc=$?; tail -3 gpurun_out/s3_bench_n1.err; grep -E "Elapsed|Maximum resident" gpurun_out/s3_bench_time.txt
echo "== bench value patterns (opt-in)"; timeout 600 python bench.py --flags 0x1000000 --no-other-configs --no-cpu-baseline > gpurun_out/s3_bench_value_patterns.json 2> gpurun_out/s3_bench_vp.err; echo rc=$?; tail -3 gpurun_out/s3_bench_vp.err
echo "== host csrspmv synthetic"; LC_ALL=C timeout 300 ellspmv_b200/host/bin/csrspmv -v --synthetic=laplace2d:8192,8192 --repeat=5 --warmup=2 -q 2>&1 | tail -9
LC_ALL=C timeout 300 ellspmv_b200/host/bin/csrspmv64 -v --synthetic=stencil27:384,384,384 --repeat=5 --warmup=2 -q 2>&1 | tail -9
du -sh gpurun_out
