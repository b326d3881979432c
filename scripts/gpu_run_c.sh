set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; tail -5 gpurun_out/r02c_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02c_bench_384.json 2> gpurun_out/r02c_bench_384.err; tail -2 gpurun_out/r02c_bench_384.err
GYSK_MERGE_SMEM_N=512 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02c_bench_512.json 2> gpurun_out/r02c_bench_512.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ingest_kernel|os_pass|runs_mark|runs_sum|bins_merge" -c 64 --launch-skip 40 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02c_ncu_bench.log 2>&1
echo done
