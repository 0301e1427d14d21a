# GPU visit: ncu --set full source-level capture of one launch of each kernel named in $KERNELS (regex list)
mkdir -p gpurun_out
TAG=${1:-p}
export OPENCV_LOG_LEVEL=ERROR
for k in $KERNELS; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$k" -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_$k python bench.py --steps 2 --warmup 3 --row-groups 2 --skip-cpu-baseline --skip-cold > gpurun_out/prof_${TAG}_$k.log 2>&1; echo "prof $k rc=$?"
done
ls -la gpurun_out/prof_${TAG}_*.ncu-rep
