mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 50 --warmup 5 --no-config5 --e2e-steps 1 ${EXTRA} > gpurun_out/sync_$name.json 2> gpurun_out/sync_$name.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/sync_$name.json').read().strip().splitlines()[-1])
print('$name', d['ms_per_step'], d['parity_check']['bit_equal'], d['exchange']['barrier'])"; }
EXTRA="--barrier fused" run fused_poll100 A=1
EXTRA="--barrier fused" run fused_poll32 ELLSPMV_CUDA_SYNC_POLL_NS=32
EXTRA="--barrier fused" run fused_poll1000 ELLSPMV_CUDA_SYNC_POLL_NS=1000
EXTRA="--barrier fused" run fused_nowait ELLSPMV_CUDA_SYNC_NOWAIT=1
EXTRA="--barrier device" run device A=1
EXTRA="--barrier nccl" run nccl A=1
