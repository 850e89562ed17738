set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "occupancy" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_facade.py -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | tail -1 > gpurun_out/r02q_bench_noextra.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench_noextra.json').read())
print(d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity_sample'])
PY
