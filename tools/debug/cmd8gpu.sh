mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_group.py tests/test_separate_diagonal.py::test_gpu_separate_diagonal_on_several_gpus tests/test_gpu_ell.py::test_exchange_on_one_gpu -m gpu -x -q --timeout 300 2>&1 | tee gpurun_out/r2_multigpu_pytest.log | tail -4
for n in 8 4; do
echo "== bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo rc=$?; tail -c 600 gpurun_out/r2_bench_n$n.err | tail -3
done
echo "== reference arm N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_n8.json 2> gpurun_out/r2_bench_ref_n8.err; echo rc=$?
echo "== C program, 8 GPUs, config 5"; timeout 300 ellspmv_b200/host/bin/ellspmv64 --gpus=8 --synthetic=stencil27s:768,768,768 --iterate --repeat=20 --warmup=3 -v -q 2>&1 | tail -2
