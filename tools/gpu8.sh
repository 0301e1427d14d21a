mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout=120 -x 2>&1 | tail -5
timeout 600 python bench.py --steps 16 --warmup 8 --skip-cpu-baseline > gpurun_out/bench_b200_v4.json 2> gpurun_out/bench_b200_v4.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b200_v4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_b200_v4.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e'])"
