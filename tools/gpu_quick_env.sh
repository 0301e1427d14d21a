# GPU visit: decode parity tests, then the C2 kernel timings once per environment assignment in $PST_ENVS (space separated,
# "-" = none), e.g. PST_ENVS="- PST_IDX_CLUSTER=0"
mkdir -p gpurun_out
TAG=${1:-q}
export OPENCV_LOG_LEVEL=ERROR
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout=300 -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$TAG.log
i=0
for e in ${PST_ENVS:--}; do
i=$((i+1))
if [ "$e" = "-" ]; then run="env"; else run="env $e"; fi
timeout 600 $run python bench.py --steps 32 --warmup 4 --skip-cpu-baseline --skip-cold > gpurun_out/bench_c2_${TAG}_$i.json 2> gpurun_out/bench_c2_${TAG}_$i.err; echo "bench c2 [$e] rc=$?"; tail -2 gpurun_out/bench_c2_${TAG}_$i.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_${TAG}_$i.json'))
print('[$e] c2 value %.4g e2e %.4g launches %s' % (d['value'], d['e2e']['value'], d.get('gpu_launches')))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
done
