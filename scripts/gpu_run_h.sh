set -x
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 300 gpurun_out/r02h_bench.err
timeout 400 python bench.py --impl reference > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err; tail -c 300 gpurun_out/r02h_bench_ref.err
K='regex:ingest_kernel|os_pass|runs_mark|runs_sum|bins_merge'
timeout 600 ncu --set full --import-source on --clock-control none -k "$K" --launch-skip 40 --launch-count 8 -f -o gpurun_out/r02h_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02h_ncu.log 2>&1
ncu -i gpurun_out/r02h_full.ncu-rep --page raw --csv > gpurun_out/r02h_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02h_full.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 64 --launch-skip 40 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02h_ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; tail -2 gpurun_out/r02h_smoke.log
echo done
