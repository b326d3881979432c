set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 scripts/sustained_stream.py > gpurun_out/r02l_sustained_n4.out 2>&1; grep '^{"config' gpurun_out/r02l_sustained_n4.out > gpurun_out/r02l_sustained_n4.json; tail -c 1200 gpurun_out/r02l_sustained_n4.json; tail -4 gpurun_out/r02l_sustained_n4.out | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 tests/run_config3.py > gpurun_out/r02l_config3_n4.out 2>&1; grep '^{"config' gpurun_out/r02l_config3_n4.out > gpurun_out/r02l_config3_n4.json; head -c 900 gpurun_out/r02l_config3_n4.json
echo done
