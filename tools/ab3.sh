set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 50 --warmup 3 > gpurun_out/ab_cvt1.json 2> gpurun_out/ab_cvt1.err
python - <<'PY'
import json
for f in ['gpurun_out/ab_cvt1.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['e2e'], d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cvt.csv \
    python bench.py --workload h1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cvt_stdout.log 2>&1
