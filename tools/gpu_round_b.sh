# GPU visit: parity tests + the four make_reader workloads of bench.py (short runs)
mkdir -p gpurun_out
TAG=${1:-r2b}
export OPENCV_LOG_LEVEL=ERROR
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_$TAG.log
for w in c1 c3 c4 c5; do
  timeout 900 python bench.py --workload $w --steps 16 --warmup 4 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"; tail -3 gpurun_out/bench_${w}_$TAG.err; cut -c1-1800 gpurun_out/bench_${w}_$TAG.json
done
