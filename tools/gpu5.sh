mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
python -m pytest tests/test_gpu_readers.py tests/test_gpu_loaders.py -m gpu -q --timeout=900 2>&1 | tail -60 > gpurun_out/t5.log; tail -8 gpurun_out/t5.log
ncu --set full --clock-control none --import-source on -k regex:'k_snappy_pages' -s 4 -c 1 -o gpurun_out/prof_r1_snappy_v2 python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_v2.log 2>&1; echo "ncu rc=$?"
