mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ell.py -m gpu -x -q -k "pattern or golden or lane" > gpurun_out/lanes_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/lanes_pytest.log
for i in 1 2; do timeout 900 python tools/bench_configs.py --configs c3,c2 --variants auto --no-csr >> gpurun_out/r2_lane_patterns.jsonl 2>gpurun_out/lanes_bc.err || tail -5 gpurun_out/lanes_bc.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r2_lane_patterns.jsonl'):
    d=json.loads(l); print(d.get('config'), d.get('variant'), d.get('mode'), d.get('ms_median'), d.get('pattern_rows_frac'), d.get('gbs_as_stored'))
PY
