# GPU visit: parity tests + the four make_reader workloads + their ncu launch lists
mkdir -p gpurun_out
TAG=${1:-r2c}
export OPENCV_LOG_LEVEL=ERROR
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/t_$TAG.log
for w in c1 c3 c4 c5; do
  timeout 900 python bench.py --workload $w --steps 16 --warmup 4 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"; tail -3 gpurun_out/bench_${w}_$TAG.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${w}_$TAG.json'))
print('$w value %.4g e2e %.4g cpu %.4g' % (d['value'], d['e2e']['value'], (d.get('cpu_baseline') or {}).get('value', 0)))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'],4)) for k,v in d['roofline']['per_kernel'].items()}))
print(json.dumps(d['e2e'])[:600])
PY
done
for w in c1 c3 c4 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${w}_$TAG.csv python bench.py --workload $w --steps 2 --warmup 3 --row-groups 4 --skip-cpu-baseline > gpurun_out/ncu_l_${w}_$TAG.log 2>&1; echo "launch list $w rc=$?"
done
