set -x
mkdir -p gpurun_out
WL=${1:-t2d}; N=${2:-512}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:patch_kernel -s 4 -c 1 -f -o gpurun_out/prof_patch_$WL \
    python bench.py --workload $WL --size $N --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_patch_$WL.log 2>&1
tail -3 gpurun_out/ncu_patch_$WL.log
ls -la gpurun_out/*.ncu-rep
