set -x
timeout 700 python bench.py --workload v2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/v2_pair1.json 2> gpurun_out/v2_pair1.err
OC_GEMM_PAIR=0 timeout 700 python bench.py --workload v2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/v2_pair0.json 2> gpurun_out/v2_pair0.err
python - <<'PY'
import json
for f in ['gpurun_out/v2_pair1.json','gpurun_out/v2_pair0.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], d['stage_ms_per_step'], d.get('roofline_tensor'), d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
