mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_group.py tests/test_separate_diagonal.py tests/test_gpu_ell.py::test_exchange_on_one_gpu -m gpu -x -q --timeout 300 > gpurun_out/r2_multigpu_pytest_n2_session3.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2_multigpu_pytest_n2_session3.log
echo "== bench N=2"; T0=$(date +%s); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo rc=$? wall=$(( $(date +%s) - T0 ))s; tail -c 600 gpurun_out/r2_bench_n2.err | tail -3
echo "== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err; echo rc=$?
echo "== numa"; (lscpu | grep -i numa; nvidia-smi topo -m) > gpurun_out/n2_topo.txt 2>&1; head -20 gpurun_out/n2_topo.txt
