# A/B of the batched-scan sweep variants on one B200 (run under gpurun): the GEMM tests, then the
# default bench per variant.  OC_GEMM_CVT=0 -> tf32 CTA pairs; OC_GEMM_PAIR=0 -> two query groups per CTA.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/ab_cvt.json 2> gpurun_out/ab_cvt.err
OC_GEMM_CVT=0 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tf32_pair.json 2> gpurun_out/ab_tf32_pair.err
OC_GEMM_CVT=0 OC_GEMM_PAIR=0 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tf32_ng2.json 2> gpurun_out/ab_tf32_ng2.err
python - <<'PY'
import json
for f in ['ab_cvt', 'ab_tf32_pair', 'ab_tf32_ng2']:
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['scan'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/{f}.err').read()[-2000:])
PY
