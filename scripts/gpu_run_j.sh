set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; tail -6 gpurun_out/r02j_tests.log
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 python bench.py $B > gpurun_out/r02j_side1.json 2> gpurun_out/r02j_side1.err
GYSK_SIDE_DRAIN=0 timeout 200 python bench.py $B > gpurun_out/r02j_side0.json 2> gpurun_out/r02j_side0.err
GYSK_SIDE_CTAS=2 timeout 200 python bench.py $B > gpurun_out/r02j_side2.json 2> gpurun_out/r02j_side2.err
GYSK_INGEST_VARIANT=842 timeout 200 python bench.py $B > gpurun_out/r02j_side1_842.json 2> gpurun_out/r02j_side1_842.err
K='regex:ingest_kernel|os_pass|runs_mark|runs_sum|bins_merge|side_drain'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 36 --launch-skip 45 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02j_ncu_bench.log 2>&1
echo done
