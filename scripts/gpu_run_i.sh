set -x
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for a in 0 1 2 4 8 15; do GYSK_EXP_ABLATE=$a timeout 200 python bench.py $B > gpurun_out/r02i_ablate_$a.json 2> gpurun_out/r02i_ablate_$a.err; done
echo done
