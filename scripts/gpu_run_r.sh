set -x
mkdir -p gpurun_out
GYSK_HOT_MIN=64 timeout 200 compute-sanitizer --tool memcheck python scripts/sanitizer_workload.py > gpurun_out/r02r_memcheck_d8.log 2>&1; tail -3 gpurun_out/r02r_memcheck_d8.log
GYSK_HOT_MIN=64 timeout 330 compute-sanitizer --tool racecheck python scripts/sanitizer_workload.py > gpurun_out/r02r_racecheck_d8.log 2>&1; tail -3 gpurun_out/r02r_racecheck_d8.log
echo done
