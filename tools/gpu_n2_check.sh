# 2-GPU visit: C3 (four row-groups per rank: the `value` leg must scale like `e2e`) and C2, one rank per GPU over NCCL
mkdir -p gpurun_out
TAG=${1:-r2n2}
export OPENCV_LOG_LEVEL=ERROR
N=2
run() {  # workload, extra args
  w=$1; shift
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w "$@" > gpurun_out/bench_${w}_n${N}_$TAG.json 2> gpurun_out/bench_${w}_n${N}_$TAG.err
  echo "bench $w N=$N rc=$?"; tail -2 gpurun_out/bench_${w}_n${N}_$TAG.err | cut -c1-300
}
run c3 --steps 16 --warmup 4 --skip-cpu-baseline
run c2 --steps 32 --warmup 4 --skip-cold --skip-cpu-baseline
python - <<PY
import json
for w in ('c3','c2'):
    try:
        d=json.loads(open('gpurun_out/bench_%s_n${N}_$TAG.json' % w).read().strip().splitlines()[-1])
        print('%s N=%d value %.4g e2e %.4g h2d %.3g GB/s' % (w, d['n_gpus'], d['value'], d['e2e']['value'], d['e2e'].get('h2d_gbps', 0)))
    except Exception as e:
        print(w, 'FAILED', e)
PY
