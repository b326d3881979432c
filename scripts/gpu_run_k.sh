set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_tests.log 2>&1; tail -6 gpurun_out/r02k_tests.log
timeout 400 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -3 gpurun_out/r02k_bench.err; head -c 600 gpurun_out/r02k_bench.json
echo done
