# GPU visit: ncu --set full source-level capture of one k_snappy_pages launch per library variant
mkdir -p gpurun_out
TAG=${1:-p}
export OPENCV_LOG_LEVEL=ERROR
for v in main $PST_VARIANTS; do
if [ $v = main ]; then unset PST_B200_LIB; else export PST_B200_LIB=$PWD/petastorm_b200/variants/libpst_$v.so; fi
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_snappy_pages' -s 8 -c 1 -f -o gpurun_out/prof_${TAG}_$v python bench.py --steps 2 --warmup 3 --row-groups 2 --skip-cpu-baseline --skip-cold > gpurun_out/prof_${TAG}_$v.log 2>&1; echo "prof $v rc=$?"
done
ls -la gpurun_out/prof_${TAG}_*.ncu-rep
