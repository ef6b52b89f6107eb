for v in 3 5 6 7 8; do
export B200_SPMV_VARIANT=$v
python bench.py --size 768 --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v T2D spmv',d['spmv'])"
python bench.py --workload t3d --size 40 --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v T3D spmv',d['spmv'])"
done
