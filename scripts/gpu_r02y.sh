# single GPU: multigrid cycle variants (W-cycle below the finest level, over-correction of the aggregation levels, MIS(1) aggregates)
mkdir -p gpurun_out
run() { echo "== $*"; n=$1; shift; env "$@" B200_VERBOSE=1 python scripts/pc_scale.py --n $n 2>gpurun_out/r02y.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n'], d['solve1']['its'], round(d['solve1']['solve_ms'], 1), d['solve1']['converged'])"; grep "amg hierarchy" gpurun_out/r02y.err | head -1; }
run 92 B200_AMG_MIS=1 B200_AMG_OVERCORRECT=1.5 B200_AMG_GAMMA=2
run 92 B200_AMG_MIS=1 B200_AMG_OVERCORRECT=1.75 B200_AMG_GAMMA=2
run 92 B200_AMG_MIS=2 B200_AMG_OVERCORRECT=2 B200_AMG_GAMMA=2
run 92 B200_AMG_MIS=1 B200_AMG_OVERCORRECT=2 B200_AMG_GAMMA=3
run 92 B200_AMG_MIS=1 B200_AMG_OVERCORRECT=2.25 B200_AMG_GAMMA=2
run 48 B200_AMG_MIS=1 B200_AMG_OVERCORRECT=2 B200_AMG_GAMMA=2
