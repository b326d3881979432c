set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hot_rows.py tests/test_gpu_parity.py -x -q > gpurun_out/r02n_tests.log 2>&1; tail -4 gpurun_out/r02n_tests.log
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 200 python bench.py $B > gpurun_out/r02n_$name.json 2> gpurun_out/r02n_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02n_$name.json').read().strip().splitlines()[-1])
    print('RESULT $name', round(d['value']/1e9,2), round(d['ms_per_step'],3), 'ingest', round(d['roofline']['ms_per_launch'],3), 'chain', round(d['roofline_other'][0]['ms_per_launch'],3))
except Exception as e: print('RESULT $name', 'ERR', e)
PY
}
run default X=1
run max1m GYSK_HOT_MAX=1000000
run binmax20k GYSK_HOT_BIN_MAX=20000
run rows4096 GYSK_HOT_ROWS=4096 GYSK_HOT_MIN=2048
run rows8192 GYSK_HOT_ROWS=8192 GYSK_HOT_MIN=1024
run off GYSK_HOT_ROWS=0
K='regex:ingest_kernel|os_pass|runs_mark|runs_sum|bins_merge'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 24 --launch-skip 40 --csv --log-file gpurun_out/r02n_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02n_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k "$K" --launch-skip 40 --launch-count 8 -f -o gpurun_out/r02n_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02n_ncu.log 2>&1
ls -la gpurun_out/r02n_full.ncu-rep
ncu -i gpurun_out/r02n_full.ncu-rep --page raw --csv > gpurun_out/r02n_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r02n_full.ncu-rep --page source --csv > gpurun_out/r02n_full_source.csv 2>/dev/null
gzip -f gpurun_out/r02n_full_source.csv
[ $(stat -c %s gpurun_out/r02n_full.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/r02n_full.ncu-rep
echo done
