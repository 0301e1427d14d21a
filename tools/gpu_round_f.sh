# GPU visit: old pipeline kernel vs wide kernel v2, index with 6 rounds, full tests
mkdir -p gpurun_out
TAG=${1:-r2f}
export OPENCV_LOG_LEVEL=ERROR
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_$TAG.log
for mode in wide pipeline; do
  if [ $mode = pipeline ]; then export PST_SNAPPY_PIPELINE=1; else unset PST_SNAPPY_PIPELINE; fi
  timeout 600 python bench.py --steps 16 --warmup 4 --skip-cold --skip-cpu-baseline > gpurun_out/bench_c2_${mode}_$TAG.json 2> gpurun_out/bench_c2_${mode}_$TAG.err; echo "bench c2 $mode rc=$?"; tail -3 gpurun_out/bench_c2_${mode}_$TAG.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_${mode}_$TAG.json'))
print('$mode c2 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
  timeout 600 python bench.py --workload c5 --steps 8 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_c5_${mode}_$TAG.json 2> gpurun_out/bench_c5_${mode}_$TAG.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_c5_${mode}_$TAG.json'))
print('$mode c5 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'],4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
done
unset PST_SNAPPY_PIPELINE
timeout 600 python bench.py --workload c4 --steps 16 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c4_$TAG.json'))
print('c4 value %.4g e2e %.4g h2d %.3g GB/s' % (d['value'], d['e2e']['value'], d['e2e']['h2d_gbps']))
PY
