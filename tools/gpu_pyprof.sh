# GPU visit: where does the host time of a row-reader workload go?  Two short runs per workload in $WORKLOADS: cProfile of the
# consumer thread (bench main), then of the pool's issuing thread (Python >= 3.12 allows one profiler per process).
mkdir -p gpurun_out
TAG=${1:-pp}
export OPENCV_LOG_LEVEL=ERROR
for w in ${WORKLOADS:-c5}; do
rm -f gpurun_out/pool_${w}_$TAG.prof.* gpurun_out/main_${w}_$TAG.prof
PST_BENCH_PROFILE=$PWD/gpurun_out/main_${w}_$TAG.prof timeout 150 python bench.py --workload $w --steps 12 --warmup 3 --skip-cpu-baseline > gpurun_out/pyprof_${w}_${TAG}_main.json 2> gpurun_out/pyprof_${w}_${TAG}_main.err; echo "$w main rc=$?"
PST_POOL_PROFILE=$PWD/gpurun_out/pool_${w}_$TAG.prof timeout 150 python bench.py --workload $w --steps 12 --warmup 3 --skip-cpu-baseline > gpurun_out/pyprof_${w}_${TAG}_pool.json 2> gpurun_out/pyprof_${w}_${TAG}_pool.err; echo "$w pool rc=$?"
ls -la gpurun_out/*_${w}_$TAG.prof* 2>/dev/null
done
