# GPU visit: full parity tests, then row-reader benches: "$RUNS" = space separated workload[:ENV=VAL] items
mkdir -p gpurun_out
TAG=${1:-rows}
export OPENCV_LOG_LEVEL=ERROR
timeout 300 python -m pytest tests -m gpu -q --timeout=120 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$TAG.log
i=0
for item in $RUNS; do
i=$((i+1))
w=${item%%:*}; e=${item#*:}; if [ "$e" = "$item" ]; then run="env"; else run="env $e"; fi
timeout 150 $run python bench.py --workload $w --steps 32 --warmup 4 --skip-cpu-baseline > gpurun_out/bench_${w}_${TAG}_$i.json 2> gpurun_out/bench_${w}_${TAG}_$i.err; echo "bench $item rc=$?"; tail -1 gpurun_out/bench_${w}_${TAG}_$i.err | cut -c1-200
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${w}_${TAG}_$i.json'))
    print('$item value %.4g e2e %.4g ms/step %.2f e2e ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
except Exception as ex: print('$item FAILED', ex)
PY
done
