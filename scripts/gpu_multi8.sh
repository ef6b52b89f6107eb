set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_t3d92_n$N.json | cut -c1-300
tail -3 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus $N --workload t2d --steps 10 --warmup 3 2>>gpurun_out/bench_n$N.err | tee gpurun_out/bench_t2d1024_n$N.json | cut -c1-300
