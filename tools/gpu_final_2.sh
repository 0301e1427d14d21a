# final single-GPU visit of round 2 (second half): what the driver runs at round end + all five workloads + ncu evidence of the
# decode kernels.  Tight timeouts: a hung step must not eat the budget.
mkdir -p gpurun_out
TAG=${1:-r2end}
export OPENCV_LOG_LEVEL=ERROR
timeout 420 python -m pytest tests -m gpu -q --timeout=120 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"; tail -2 gpurun_out/bench_c2_$TAG.err
for w in c1 c3 c4 c5; do
  timeout 300 python bench.py --workload $w --steps 32 --warmup 4 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"; tail -2 gpurun_out/bench_${w}_$TAG.err
done
python - <<PY
import json
for w in ('c2','c1','c3','c4','c5'):
    try:
        d=json.load(open('gpurun_out/bench_%s_$TAG.json' % w))
    except Exception as e:
        print(w, 'FAILED', e); continue
    cpu=(d.get('cpu_baseline') or {}).get('value') or 0
    print('%s value %.4g e2e %.4g cpu %.4g  e2e/cpu %.1f  dom %s frac %.3f' % (w, d['value'], d['e2e']['value'], cpu, d['e2e']['value']/cpu if cpu else 0, d['roofline']['kernel'][:30], d['roofline']['frac']))
    print('   ', json.dumps({k:(round(v['ms'],3), round(v['frac'] or 0,4)) for k,v in d['roofline']['per_kernel'].items()}))
PY
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_snappy_pages|k_snappy_index|k_decode_pages|k_copy_tiles' -s 16 -c 4 -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline --skip-cold > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --row-groups 2 --skip-cpu-baseline --skip-cold > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
