set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hot_rows.py -x -q > gpurun_out/r02l_hot_tests.log 2>&1; tail -6 gpurun_out/r02l_hot_tests.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_hot_rows.py > gpurun_out/r02l_tests.log 2>&1; tail -6 gpurun_out/r02l_tests.log
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 python bench.py $B > gpurun_out/r02l_hot_default.json 2> gpurun_out/r02l_hot_default.err
GYSK_HOT_ROWS=0 timeout 200 python bench.py $B > gpurun_out/r02l_hot_off.json 2> gpurun_out/r02l_hot_off.err
GYSK_HOT_ROWS=0 GYSK_OS_PERSIST=0 timeout 200 python bench.py $B > gpurun_out/r02l_hot_off_nopersist.json 2> gpurun_out/r02l_hot_off_nopersist.err
GYSK_OS_PERSIST=0 timeout 200 python bench.py $B > gpurun_out/r02l_hot_nopersist.json 2> gpurun_out/r02l_hot_nopersist.err
GYSK_HOT_ROWS=4096 GYSK_HOT_MIN=2048 timeout 200 python bench.py $B > gpurun_out/r02l_hot_4096.json 2> gpurun_out/r02l_hot_4096.err
GYSK_HOT_ROWS=8192 GYSK_HOT_MIN=1024 timeout 200 python bench.py $B > gpurun_out/r02l_hot_8192.json 2> gpurun_out/r02l_hot_8192.err
for f in default off off_nopersist nopersist 4096 8192; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02l_hot_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']/1e9,2), d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline_other'][0]['ms_per_launch'], d['accuracy']['max_rel_err_p99'])
except Exception as e: print('$f', 'ERR', e)
PY
done
K='regex:ingest_kernel|os_pass|runs_mark|runs_sum|bins_merge'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 24 --launch-skip 40 --csv --log-file gpurun_out/r02l_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02l_ncu_bench.log 2>&1
tail -12 gpurun_out/r02l_launches.csv | cut -c1-200
echo done
