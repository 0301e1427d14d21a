# one --set full capture of the Snappy kernel (and the page decoder) + the launch list of a short bench run
mkdir -p gpurun_out
TAG=${1:-r1_snappy_v5}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_snappy_pages|k_snappy_index|k_decode_pages' -s 20 -c 4 -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --row-groups 2 --skip-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
