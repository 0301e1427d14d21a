# everything the driver runs at round end, on one GPU
mkdir -p gpurun_out
export OPENCV_LOG_LEVEL=ERROR
timeout 900 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -2 gpurun_out/bench_reference.err
cat gpurun_out/bench_default.json | cut -c1-3000
cat gpurun_out/bench_reference.json | cut -c1-1200
