# GPU visit: C2 bench with the cold e2e leg, once per environment assignment in $PST_ENVS ("-" = none)
mkdir -p gpurun_out
TAG=${1:-cold}
export OPENCV_LOG_LEVEL=ERROR
i=0
for e in ${PST_ENVS:--}; do
i=$((i+1))
if [ "$e" = "-" ]; then run="env"; else run="env $e"; fi
timeout 200 $run python bench.py --steps 32 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_c2_${TAG}_$i.json 2> gpurun_out/bench_c2_${TAG}_$i.err; echo "bench c2 [$e] rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_${TAG}_$i.json'))
c=d['e2e']['cold']
print('[$e] value %.4g e2e %.4g (%.2f ms) cold %.4g (%.2f ms, %.1f GB/s)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], c['value'], c['ms_per_step'], c['h2d_gbps']))
PY
done
