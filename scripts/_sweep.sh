for c in 42 82 44 24 84 162; do echo "== B200_AMG_SPMV=$c"; B200_AMG_SPMV=$c python scripts/pc_scale.py --n 64 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['solve1']['its'], d['solve1']['solve_ms'], d['solve1']['solve_ms']/d['solve1']['its'])"; done
