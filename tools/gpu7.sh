mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout=120 -x 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_snappy_pages|k_decode_pages' -s 8 -c 2 -o gpurun_out/prof_r1_snappy_v3 python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_v3.log 2>&1; echo "ncu rc=$?"
