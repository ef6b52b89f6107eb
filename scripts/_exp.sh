set -x
python -m pytest tests/test_gpu_precond.py -q 2>&1 | tail -3
for env in "B200_AMG_F32=0" "B200_AMG_F32=1" "B200_AMG_DEGREE=3" "B200_AMG_CYCLES=2" "B200_AMG_DEGREE=4" "B200_AMG_RATIO=8" "B200_AMG_DEGREE=3 B200_AMG_RATIO=10"; do
  echo "== $env"; env $env python scripts/pc_scale.py --n 48 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['solve1'])"
done
for r in 60 100; do echo "== restart $r"; python scripts/pc_scale.py --n 48 --restart $r 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['solve1'])"; done
echo "== N=92 default"; python scripts/pc_scale.py --n 92 2>&1 | tail -1
echo "== N=92 deg3"; B200_AMG_DEGREE=3 python scripts/pc_scale.py --n 92 2>&1 | tail -1
