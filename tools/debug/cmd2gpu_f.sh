# split-launch exchange, three modes: 2 = boundary and interior side by side (default), 1 = in sequence, 0 = one launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_group.py tests/test_separate_diagonal.py "tests/test_gpu_ell.py::test_exchange_on_one_gpu" -m gpu -x -q > gpurun_out/split_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/split_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 50 --warmup 5 --e2e-steps 1 --no-other-configs --config5-iters 30 > gpurun_out/split_$name.json 2> gpurun_out/split_$name.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/split_$name.json').read().strip().splitlines()[-1])
c5=(d.get('config5') or {}).get('push') or {}
print('$name', d['ms_per_step'], d['parity_check']['bit_equal'], (d.get('scaling_base') or {}), c5.get('ms_per_step'), (c5.get('parity_check') or {}).get('bit_equal'))" || tail -5 gpurun_out/split_$name.err; }
run m2 ELLSPMV_CUDA_SPLIT_EXCHANGE=2
run m1 ELLSPMV_CUDA_SPLIT_EXCHANGE=1
run m0 ELLSPMV_CUDA_SPLIT_EXCHANGE=0
run m2b ELLSPMV_CUDA_SPLIT_EXCHANGE=2
run m1b ELLSPMV_CUDA_SPLIT_EXCHANGE=1
run m0b ELLSPMV_CUDA_SPLIT_EXCHANGE=0
