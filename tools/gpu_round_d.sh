# GPU visit: parity tests with the wide Snappy fragment kernel, the warp-cooperative PNG inflate and device-bitstream JPEG
mkdir -p gpurun_out
TAG=${1:-r2d}
export OPENCV_LOG_LEVEL=ERROR
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/t_$TAG.log
timeout 600 python bench.py --steps 16 --warmup 4 --skip-cold > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"; tail -3 gpurun_out/bench_c2_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_$TAG.json'))
print('c2 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
for w in c1 c3 c5; do
  timeout 900 python bench.py --workload $w --steps 8 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"; tail -3 gpurun_out/bench_${w}_$TAG.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${w}_$TAG.json'))
print('$w value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'],4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
done
python -c "
from petastorm_b200 import device_ops
print('jpeg device backend', device_ops.jpeg_device_backend())"
