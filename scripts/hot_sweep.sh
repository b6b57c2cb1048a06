for MODE in 1 2 3; do
PGLB_HOT_MODE=$MODE PGLB_HOT_MB=48 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:stream128 -s 4 -c 1 --csv --log-file gpurun_out/hot_m$MODE.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
echo MODE $MODE; grep -E "dram__bytes_read|hit_rate|time_duration" gpurun_out/hot_m$MODE.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
PGLB_HOT_MB=0 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:stream128 -s 4 -c 1 --csv --log-file gpurun_out/hot_off.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
echo OFF; grep -E "dram__bytes_read|hit_rate|time_duration" gpurun_out/hot_off.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
for MODE in 2 3; do PGLB_HOT_MODE=$MODE PGLB_HOT_MB=48 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MODE', $MODE, d['ms_per_step'], d['roofline']['frac'])"; done
