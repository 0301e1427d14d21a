# one GPU visit: parity tests, a short C2 bench, ncu launch list + one --set full capture of the decode kernels
mkdir -p gpurun_out
TAG=${1:-r2a}
export OPENCV_LOG_LEVEL=ERROR
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=300 > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_$TAG.log
timeout 900 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_snappy_pages|k_snappy_index|k_decode_pages|k_copy_tiles' -s 16 -c 4 -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
cut -c1-2500 gpurun_out/bench_$TAG.json
