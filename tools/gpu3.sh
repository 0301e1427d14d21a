mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" 
python -m pytest tests/test_gpu_readers.py -m gpu -q --timeout=900 2>&1 | tail -40 > gpurun_out/t3.log; tail -5 gpurun_out/t3.log
( time python bench.py --steps 16 --warmup 8 ) > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b200.err
cat gpurun_out/bench_b200.json | cut -c1-3000
( time python bench.py --impl reference --steps 16 --warmup 8 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json | cut -c1-1500
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_snappy_pages|k_decode_pages' -s 6 -c 2 -o gpurun_out/prof_r1 python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
