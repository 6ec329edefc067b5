# experiment: L2 bulk prefetch distance (slices ahead) for the default ELL kernel
set -x
out=gpurun_out/r1_prefetch_sweep2.jsonl
: > $out
for d in 64 128 200 300 400 600; do
  echo "# prefetch_slices=$d" >> $out
  ELLSPMV_CUDA_PREFETCH_SLICES=$d timeout 120 python tools/bench_configs.py --configs c2,c3 --variants auto --no-csr --reps 20 2>> gpurun_out/r1_prefetch_sweep2.err | grep accumulate | cut -c1-230 >> $out
done
for d in 128 300; do
  echo "# no_pattern prefetch_slices=$d" >> $out
  ELLSPMV_CUDA_PREFETCH_SLICES=$d timeout 120 python tools/bench_configs.py --configs c2,c3 --variants auto --no-csr --reps 20 --extra-flags 0x40000 2>> gpurun_out/r1_prefetch_sweep2.err | grep accumulate | cut -c1-230 >> $out
done
echo "# c4 prefetch_slices=300" >> $out
ELLSPMV_CUDA_PREFETCH_SLICES=300 timeout 120 python tools/bench_configs.py --configs c4 --variants auto,staged --no-csr --reps 10 2>> gpurun_out/r1_prefetch_sweep2.err | grep accumulate | cut -c1-230 >> $out
cat $out
timeout 20 python -c "import torch; torch.zeros(4,device='cuda').sum().item(); print('alive')"
