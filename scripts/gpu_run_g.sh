set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; tail -6 gpurun_out/r02g_tests.log
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 python bench.py $B > gpurun_out/r02g_v832.json 2> gpurun_out/r02g_v832.err
GYSK_INGEST_VARIANT=842 timeout 200 python bench.py $B > gpurun_out/r02g_v842.json 2> gpurun_out/r02g_v842.err
GYSK_OS_NARROW=0 timeout 200 python bench.py $B > gpurun_out/r02g_wide.json 2> gpurun_out/r02g_wide.err
GYSK_INGEST_VARIANT=842 timeout 200 python bench.py $B > gpurun_out/r02g_v842b.json 2> gpurun_out/r02g_v842b.err
timeout 200 python bench.py $B > gpurun_out/r02g_v832b.json 2> gpurun_out/r02g_v832b.err
echo done
