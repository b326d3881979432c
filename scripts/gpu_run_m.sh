set -x
mkdir -p gpurun_out
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 200 python bench.py $B > gpurun_out/r02m_$name.json 2> gpurun_out/r02m_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02m_$name.json').read().strip().splitlines()[-1])
    print('RESULT $name', round(d['value']/1e9,2), round(d['ms_per_step'],3), 'ingest', round(d['roofline']['ms_per_launch'],3), 'chain', round(d['roofline_other'][0]['ms_per_launch'],3))
except Exception as e: print('RESULT $name', 'ERR', e)
PY
}
run aos X=1
run soa GYSK_HOT_SOA=1
run onered GYSK_EXP_ABLATE=16
run max1m GYSK_HOT_MAX=1000000
run max100k GYSK_HOT_MAX=100000
run soa_max1m GYSK_HOT_SOA=1 GYSK_HOT_MAX=1000000
run min64k GYSK_HOT_MIN=65536
echo done
