# GPU visit: full parity tests + C5 / C1 benches (loader-side changes)
mkdir -p gpurun_out
TAG=${1:-ld}
export OPENCV_LOG_LEVEL=ERROR
timeout 300 python -m pytest tests -m gpu -q --timeout=120 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$TAG.log
for w in ${WORKLOADS:-c5}; do
timeout 200 python bench.py --workload $w --steps 32 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"; tail -2 gpurun_out/bench_${w}_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${w}_$TAG.json'))
print('$w value %.4g e2e %.4g ms/step %.2f e2e ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done
