# one ncu --set full capture (with source) of the wide Snappy fragment kernel on a C2 row-group + the launch list
mkdir -p gpurun_out
TAG=${1:-r2e}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_snappy_frag|k_snappy_index|k_decode_pages' -s 12 -c 3 -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline --skip-cold > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 16 --warmup 4 --skip-cold --skip-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"; tail -3 gpurun_out/bench_c2_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_$TAG.json'))
print('c2 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_$TAG.log
