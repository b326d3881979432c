set -x
mkdir -p gpurun_out
GYSK_KEY_DIGIT_MAX=9 GYSK_HOT_MIN=64 timeout 100 compute-sanitizer --tool memcheck python scripts/sanitizer_workload.py > gpurun_out/r02s_memcheck_d9.log 2>&1; tail -3 gpurun_out/r02s_memcheck_d9.log
GYSK_KEY_DIGIT_MAX=9 GYSK_HOT_MIN=64 timeout 150 compute-sanitizer --tool racecheck python scripts/sanitizer_workload.py > gpurun_out/r02s_racecheck_d9.log 2>&1; tail -3 gpurun_out/r02s_racecheck_d9.log
echo done
