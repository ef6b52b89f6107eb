# r02y: multigrid cycle variants on one GPU (pc_scale.py solves the bench Jacobian twice; the second solve is reported)
#   knobs: B200_AMG_MIS (1 | 2), B200_AMG_GAMMA (cycle index below the finest level), B200_AMG_OVERCORRECT (aggregation levels)
# The sweep was run in three calls (profiles/README.md r02y has the table); this file keeps the full list.
mkdir -p gpurun_out
run() { echo "== $*"; n=$1; shift; env "$@" B200_VERBOSE=1 python scripts/pc_scale.py --n $n 2>gpurun_out/r02y.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], d['solve1']['its'], round(d['solve1']['solve_ms'], 1), d['solve1']['converged'])"; grep "amg hierarchy" gpurun_out/r02y.err | head -1; }
# the defaults at the time of the sweep were MIS=2 GAMMA=1 OVERCORRECT=1
for n in 48 92; do
  run $n B200_AMG_MIS=2 B200_AMG_GAMMA=1 B200_AMG_OVERCORRECT=1
  run $n B200_AMG_MIS=2 B200_AMG_GAMMA=1 B200_AMG_OVERCORRECT=1.5
  run $n B200_AMG_MIS=2 B200_AMG_GAMMA=1 B200_AMG_OVERCORRECT=2
  run $n B200_AMG_MIS=1 B200_AMG_GAMMA=1 B200_AMG_OVERCORRECT=1.5
  run $n B200_AMG_MIS=1 B200_AMG_GAMMA=2 B200_AMG_OVERCORRECT=1.5
  run $n B200_AMG_MIS=1 B200_AMG_GAMMA=2 B200_AMG_OVERCORRECT=1.75
  run $n B200_AMG_MIS=1 B200_AMG_GAMMA=2 B200_AMG_OVERCORRECT=2
done
