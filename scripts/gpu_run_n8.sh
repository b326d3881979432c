set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/sustained_stream.py > gpurun_out/r02f_sustained_n8.out 2>&1; grep '^{"config' gpurun_out/r02f_sustained_n8.out > gpurun_out/r02f_sustained_n8.json; tail -c 1500 gpurun_out/r02f_sustained_n8.json; tail -4 gpurun_out/r02f_sustained_n8.out | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_n8.out 2>&1; grep '^{"metric' gpurun_out/r02f_bench_n8.out > gpurun_out/r02f_bench_n8.json; tail -c 300 gpurun_out/r02f_bench_n8.json
echo done
