mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
python -m pytest tests/test_gpu_decode.py tests/test_gpu_ops.py tests/test_gpu_readers.py tests/test_gpu_loaders.py -m gpu -q --timeout=900 2>&1 | tail -80 > gpurun_out/t4.log; tail -30 gpurun_out/t4.log
python bench.py --steps 16 --warmup 8 --skip-cpu-baseline > gpurun_out/bench_b200_v2.json 2> gpurun_out/bench_b200_v2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b200_v2.err
cat gpurun_out/bench_b200_v2.json | cut -c1-2500
