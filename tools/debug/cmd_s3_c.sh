mkdir -p gpurun_out
echo "== pytest gpu (csr, host programs, diagonal)"; timeout 1200 python -m pytest tests/test_gpu_csr.py tests/test_gpu_host_programs.py tests/test_separate_diagonal.py -m gpu -x -q > gpurun_out/s3_pytest_gpu_csr.log 2>&1; echo rc=$?; tail -4 gpurun_out/s3_pytest_gpu_csr.log
echo "== host csrspmv synthetic"; LC_ALL=C timeout 300 ellspmv_b200/host/bin/csrspmv -v --synthetic=laplace2d:8192,8192 --repeat=3 --warmup=2 -q 2>&1 | tail -3
LC_ALL=C timeout 300 ellspmv_b200/host/bin/csrspmv64 -v --synthetic=stencil27:384,384,384 --repeat=3 --warmup=2 -q 2>&1 | tail -3
echo "== c4 block sweep"; timeout 900 python tools/c4_blocks.py 16 24 32 40 48 > gpurun_out/r2_c4_block_sweep.jsonl 2> gpurun_out/c4b.err; echo rc=$?; cat gpurun_out/r2_c4_block_sweep.jsonl; tail -3 gpurun_out/c4b.err
echo "== ncu c2 csr"; timeout 600 ncu --set full --clock-control none -k regex:ell_thread -s 4 -c 1 -o gpurun_out/r2_c2_csr python tools/profile_target.py --config c2 --path csr > gpurun_out/ncu_c2csr.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_c2csr.log
for f in gpurun_out/*.ncu-rep; do b=${f%.ncu-rep}; ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null; ncu -i $f --page details 2>/dev/null | head -c 60000 > ${b}_details.txt; rm -f $f; done
du -sh gpurun_out
