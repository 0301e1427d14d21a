mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
python -m pytest tests/test_gpu_readers.py -m gpu -q --timeout=900 2>&1 | tail -150 > gpurun_out/t2.log
tail -100 gpurun_out/t2.log
