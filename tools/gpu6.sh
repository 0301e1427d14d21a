mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout=120 -x 2>&1 | tail -40 > gpurun_out/t6.log; tail -15 gpurun_out/t6.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_readers.py tests/test_gpu_loaders.py -m gpu -q --timeout=300 2>&1 | tail -15 > gpurun_out/t6b.log; tail -6 gpurun_out/t6b.log
timeout 600 python bench.py --steps 16 --warmup 8 --skip-cpu-baseline > gpurun_out/bench_b200_v3.json 2> gpurun_out/bench_b200_v3.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b200_v3.err
cat gpurun_out/bench_b200_v3.json | cut -c1-2500
