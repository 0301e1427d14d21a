# GPU visit: k_snappy_pages with 27 KiB of shared memory (8 CTAs/SM) - parity + C2 timings
mkdir -p gpurun_out
TAG=${1:-r2h}
export OPENCV_LOG_LEVEL=ERROR
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$TAG.log
timeout 600 python bench.py --steps 32 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"; tail -2 gpurun_out/bench_c2_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c2_$TAG.json'))
print('c2 value %.4g e2e %.4g cold %.4g (%.3g GB/s)' % (d['value'], d['e2e']['value'], d['e2e']['cold']['value'], d['e2e']['cold']['h2d_gbps']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
for w in c1 c3; do
timeout 600 python bench.py --workload $w --steps 16 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${w}_$TAG.json'))
print('$w value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
print(json.dumps({k:(round(v['ms'],3), round(v['frac'],4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_snappy_pages' -s 8 -c 2 --csv --log-file gpurun_out/occ_$TAG.csv python bench.py --steps 2 --warmup 3 --row-groups 2 --skip-cpu-baseline --skip-cold > /dev/null 2>&1; grep -c . gpurun_out/occ_$TAG.csv; tail -6 gpurun_out/occ_$TAG.csv | cut -c1-300
