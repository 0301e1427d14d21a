# GPU visit: resolver threads on/off for the make_reader workloads, full tests
mkdir -p gpurun_out
TAG=${1:-r2g}
export OPENCV_LOG_LEVEL=ERROR
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_$TAG.log
for res in 2 0; do
  export PST_RESOLVERS=$res
  for w in c1 c3 c5; do
    timeout 600 python bench.py --workload $w --steps 16 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_${w}_res${res}_$TAG.json 2> gpurun_out/bench_${w}_res${res}_$TAG.err; echo "bench $w res=$res rc=$?"; tail -2 gpurun_out/bench_${w}_res${res}_$TAG.err
    python - <<PY
import json
d=json.load(open('gpurun_out/bench_${w}_res${res}_$TAG.json'))
print('$w resolvers=$res value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
PY
  done
done
unset PST_RESOLVERS
timeout 600 python bench.py --steps 32 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_$TAG.json'))
print('c2 value %.4g e2e %.4g cold %.4g' % (d['value'], d['e2e']['value'], d['e2e']['cold']['value']))
print(json.dumps(d['e2e']['cold'])[:500])
PY
