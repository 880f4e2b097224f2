# A/B of the batched-scan kernels on one B200: tests first, then the default bench per variant
set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/ab_cvt1.json 2> gpurun_out/ab_cvt1.err
OC_GEMM_CVT=0 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/ab_cvt0.json 2> gpurun_out/ab_cvt0.err
python - <<'PY'
import json
for f in ['gpurun_out/ab_cvt1.json','gpurun_out/ab_cvt0.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['scan'], d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
