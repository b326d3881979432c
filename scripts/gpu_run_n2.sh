set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r02d_nccl_tests.log 2>&1; tail -5 gpurun_out/r02d_nccl_tests.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_n2.out 2>&1; grep "^{.metric" gpurun_out/r02d_bench_n2.out > gpurun_out/r02d_bench_n2.json; tail -c 400 gpurun_out/r02d_bench_n2.json; grep -c "NCCL INFO" gpurun_out/r02d_bench_n2.err
echo done
