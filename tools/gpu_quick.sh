# GPU visit: decode parity tests + C2 kernel timings (short); extra library variants: PST_VARIANTS="copy4 ..."
mkdir -p gpurun_out
TAG=${1:-q}
export OPENCV_LOG_LEVEL=ERROR
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout=300 -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$TAG.log
for v in main $PST_VARIANTS; do
if [ $v = main ]; then unset PST_B200_LIB; else export PST_B200_LIB=$PWD/petastorm_b200/variants/libpst_$v.so; timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout=300 -x 2>&1 | tail -1; fi
timeout 600 python bench.py --steps 32 --warmup 4 --skip-cpu-baseline --skip-cold > gpurun_out/bench_c2_${TAG}_$v.json 2> gpurun_out/bench_c2_${TAG}_$v.err; echo "bench c2 $v rc=$?"; tail -2 gpurun_out/bench_c2_${TAG}_$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_${TAG}_$v.json'))
print('$v c2 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
done
