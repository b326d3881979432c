set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r02o_nccl_tests.log 2>&1; tail -3 gpurun_out/r02o_nccl_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tests/run_config3.py > gpurun_out/r02o_config3_n4.out 2>&1; grep '^{"config' gpurun_out/r02o_config3_n4.out > gpurun_out/r02o_config3_n4.json; tail -c 1500 gpurun_out/r02o_config3_n4.json; tail -5 gpurun_out/r02o_config3_n4.out | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_n4.out 2>&1; grep '^{"metric' gpurun_out/r02o_bench_n4.out > gpurun_out/r02o_bench_n4.json; head -c 400 gpurun_out/r02o_bench_n4.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_n2.out 2>&1; grep '^{"metric' gpurun_out/r02o_bench_n2.out > gpurun_out/r02o_bench_n2.json; head -c 400 gpurun_out/r02o_bench_n2.json
echo done
