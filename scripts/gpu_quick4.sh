set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gather_|element_state" -s 60 -c 40 --csv --log-file gpurun_out/lane_list_t3d.csv \
    python bench.py --workload t3d --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu4.log 2>&1
