set -x
mkdir -p gpurun_out
K='regex:ingest_kernel|os_pass|runs_mark|runs_sum|bins_merge'
timeout 600 ncu --set full --import-source on --clock-control none -k "$K" --launch-skip 40 --launch-count 8 -f -o gpurun_out/r02b_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_ncu.log 2>&1
ls -la gpurun_out/r02b_full.ncu-rep
ncu -i gpurun_out/r02b_full.ncu-rep --page raw --csv > gpurun_out/r02b_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r02b_full.ncu-rep --page source --csv > gpurun_out/r02b_full_source.csv 2>/dev/null
gzip -f gpurun_out/r02b_full_source.csv
[ $(stat -c %s gpurun_out/r02b_full.ncu-rep) -gt 45000000 ] && rm -f gpurun_out/r02b_full.ncu-rep
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitizer_workload.py > gpurun_out/r02b_memcheck_d8.log 2>&1; tail -3 gpurun_out/r02b_memcheck_d8.log
GYSK_KEY_DIGIT_MAX=9 timeout 400 compute-sanitizer --tool memcheck python scripts/sanitizer_workload.py > gpurun_out/r02b_memcheck_d9.log 2>&1; tail -3 gpurun_out/r02b_memcheck_d9.log
timeout 500 compute-sanitizer --tool racecheck python scripts/sanitizer_workload.py > gpurun_out/r02b_racecheck_d8.log 2>&1; tail -3 gpurun_out/r02b_racecheck_d8.log
GYSK_KEY_DIGIT_MAX=9 timeout 600 compute-sanitizer --tool racecheck python scripts/sanitizer_workload.py > gpurun_out/r02b_racecheck_d9.log 2>&1; tail -3 gpurun_out/r02b_racecheck_d9.log
timeout 300 python scripts/sustained_stream.py --windows 12 --events-per-window 2e8 --services 125000 > gpurun_out/r02b_sustained_sanity.json 2> gpurun_out/r02b_sustained_sanity.err; tail -c 1500 gpurun_out/r02b_sustained_sanity.json; tail -5 gpurun_out/r02b_sustained_sanity.err
timeout 300 python tests/run_config3.py --events 4e6 --batch 2000000 --services 2000 --cpu-threads 4 > gpurun_out/r02b_config3_sanity.json 2> gpurun_out/r02b_config3_sanity.err; tail -c 1500 gpurun_out/r02b_config3_sanity.json; tail -5 gpurun_out/r02b_config3_sanity.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_wire.json 2> gpurun_out/r02b_bench_wire.err; tail -3 gpurun_out/r02b_bench_wire.err
for v in 1832 1834 832; do GYSK_INGEST_VARIANT=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_variant_$v.json 2> gpurun_out/r02b_variant_$v.err; done
echo done
