# 8-GPU visit (one box): the 8-GPU configurations of BASELINE.json (C3, C4, C5) + C2, one rank per GPU over NCCL
mkdir -p gpurun_out
TAG=${1:-r2n8}
export OPENCV_LOG_LEVEL=ERROR
N=${2:-8}
run() {  # workload, extra args
  w=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w "$@" > gpurun_out/bench_${w}_n${N}_$TAG.json 2> gpurun_out/bench_${w}_n${N}_$TAG.err
  echo "bench $w N=$N rc=$?"; tail -2 gpurun_out/bench_${w}_n${N}_$TAG.err | cut -c1-300
}
run c3 --steps 16 --warmup 4
run c4 --steps 16 --warmup 4
run c5 --steps 16 --warmup 4
run c2 --steps 32 --warmup 4 --skip-cold
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 16 --warmup 4 > gpurun_out/bench_c2_ref_n${N}_$TAG.json 2> gpurun_out/bench_c2_ref_n${N}_$TAG.err; echo "ref c2 N=$N rc=$?"
python - <<PY
import json
for w in ('c3','c4','c5','c2'):
    try:
        d=json.loads(open('gpurun_out/bench_%s_n${N}_$TAG.json' % w).read().strip().splitlines()[-1])
        print('%s N=%d value %.4g e2e %.4g h2d %.3g GB/s' % (w, d['n_gpus'], d['value'], d['e2e']['value'], d['e2e'].get('h2d_gbps', 0)))
    except Exception as e:
        print(w, 'FAILED', e)
try:
    d=json.loads(open('gpurun_out/bench_c2_ref_n${N}_$TAG.json').read().strip().splitlines()[-1]); print('ref c2 under torchrun: %.4g' % d['value'], d['cpu_baseline'].get('arrow_cpu_threads'))
except Exception as e:
    print('ref FAILED', e)
PY
