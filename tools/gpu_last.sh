# last visit of the round: exactly what the driver runs at round end (tests, smoke, both bench arms), short
mkdir -p gpurun_out
TAG=${1:-r2last}
export OPENCV_LOG_LEVEL=ERROR
timeout 300 python -m pytest tests -m gpu -x -q --timeout=120 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 280 python bench.py > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_c2_$TAG.err | cut -c1-200
timeout 200 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_c2_ref_$TAG.json 2> gpurun_out/bench_c2_ref_$TAG.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_$TAG.json'))
print('value %.4g e2e %.4g cold %.4g cpu %.4g launches %s clocks %s' % (d['value'], d['e2e']['value'], d['e2e']['cold']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
r=json.load(open('gpurun_out/bench_c2_ref_$TAG.json')); print('reference arm %.4g %s' % (r['value'], r['impl']))
PY
