set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; tail -4 gpurun_out/r02q_tests.log
timeout 400 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; tail -3 gpurun_out/r02q_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print('RESULT', round(d['value']/1e9,2), round(d['ms_per_step'],3), 'ingest', round(d['roofline']['ms_per_launch'],3), d['roofline']['frac'], 'chain', round(d['roofline_other'][0]['ms_per_launch'],3), 'e2e', d['e2e']['value']/1e9, 'wire', d['e2e_wire']['value']/1e9, d['hot_rows']['share_of_events'], d['hot_rows']['rows_in_use'])
PY
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02q_bench_ref.json 2> gpurun_out/r02q_bench_ref.err; head -c 600 gpurun_out/r02q_bench_ref.json
K='regex:ingest_kernel|os_pass|runs_mark|runs_sum|bins_merge'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 24 --launch-skip 40 --csv --log-file gpurun_out/r02q_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02q_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k "$K" --launch-skip 40 --launch-count 8 -f -o gpurun_out/r02q_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02q_ncu.log 2>&1
ncu -i gpurun_out/r02q_full.ncu-rep --page raw --csv > gpurun_out/r02q_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r02q_full.ncu-rep --page source --csv > gpurun_out/r02q_full_source.csv 2>/dev/null
gzip -f gpurun_out/r02q_full_source.csv
rm -f gpurun_out/r02q_full.ncu-rep
echo done
