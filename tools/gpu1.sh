mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt; df -h /tmp >> gpurun_out/gpu.txt
python -m pytest tests/test_gpu_decode.py tests/test_gpu_ops.py -m gpu -q --timeout=600 2>&1 | tail -60 > gpurun_out/t1.log
cat gpurun_out/t1.log | tail -40
