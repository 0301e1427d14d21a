# GPU visit: where does the host time of a row-reader workload go?  cProfile of the consumer thread (bench main) and of the
# pool's issuing thread, for the workloads in $WORKLOADS
mkdir -p gpurun_out
TAG=${1:-pp}
export OPENCV_LOG_LEVEL=ERROR
for w in ${WORKLOADS:-c5}; do
rm -f gpurun_out/pool_${w}_$TAG.prof.*
PST_BENCH_PROFILE=$PWD/gpurun_out/main_${w}_$TAG.prof PST_POOL_PROFILE=$PWD/gpurun_out/pool_${w}_$TAG.prof timeout 600 python bench.py --workload $w --steps 12 --warmup 3 --skip-cpu-baseline > gpurun_out/pyprof_${w}_$TAG.json 2> gpurun_out/pyprof_${w}_$TAG.err; echo "$w rc=$?"
python - <<PY
import pstats, glob, json
try:
    d=json.load(open('gpurun_out/pyprof_${w}_$TAG.json')); print('$w value %.4g e2e %.4g ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('json', e)
for f in ['gpurun_out/main_${w}_$TAG.prof'] + sorted(glob.glob('gpurun_out/pool_${w}_$TAG.prof.*')):
    print('=====', f)
    st = pstats.Stats(f); st.sort_stats('tottime').print_stats(30)
PY
done
