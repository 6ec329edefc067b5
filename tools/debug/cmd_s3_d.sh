mkdir -p gpurun_out
echo "== bulk store lab"; timeout 300 tools/experiments/bulk_store_lab > gpurun_out/r2_bulk_store_lab.jsonl 2> gpurun_out/bsl.err; echo rc=$?; cat gpurun_out/r2_bulk_store_lab.jsonl; tail -3 gpurun_out/bsl.err
echo "== bench N=1 (default command)"; T0=$(date +%s); timeout 900 python bench.py > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench_n1.err; echo rc=$? wall=$(( $(date +%s) - T0 ))s; tail -3 gpurun_out/s3_bench_n1.err
