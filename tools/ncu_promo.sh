set -x
for promo in 256 128 none; do
OC_TMA_PROMO=$promo timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_ltcfabric.sum --clock-control none -k regex:emb_gemm_cvt_kernel -s 7 -c 1 --csv --log-file gpurun_out/promo_$promo.csv \
    python bench.py --workload h1 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -E "dram__bytes_read|gpu__time|ltcfabric" gpurun_out/promo_$promo.csv | cut -d, -f12-
done
OC_TMA_PROMO=128 timeout 300 python bench.py --steps 50 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('promo128', round(d['value']), d['stage_ms_per_step'])"
