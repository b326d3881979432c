set -x
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02t_smoke.log 2>&1; tail -2 gpurun_out/r02t_smoke.log
timeout 150 python bench.py --steps 10 > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; tail -2 gpurun_out/r02t_bench.err; head -c 900 gpurun_out/r02t_bench.json
echo done
