set -x
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/sustained_stream.py > gpurun_out/r02p_sustained_n8.out 2>&1; grep '^{"config' gpurun_out/r02p_sustained_n8.out > gpurun_out/r02p_sustained_n8.json; tail -c 1500 gpurun_out/r02p_sustained_n8.json; tail -4 gpurun_out/r02p_sustained_n8.out | cut -c1-300
echo done
