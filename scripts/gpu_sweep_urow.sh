#!/bin/bash
# sweep of the row-lane kernel knobs at T3D(92): Melem/s and ms per pass
for cfg in "4 5" "4 2" "4 10" "8 5" "8 2" "8 10"; do
  set -- $cfg
  B200_UROW_NGRP=$1 B200_UROW_NB=$2 python bench.py --steps 5 --warmup 3 --no-cpu --no-solve --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NGRP=$1 NB=$2', round(d['value'],1), round(d['ms_per_step'],3))"
done
for cell in 4096 65536; do
  B200_UROW_CELL=$cell python bench.py --steps 5 --warmup 3 --no-cpu --no-solve --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CELL=$cell', round(d['value'],1), round(d['ms_per_step'],3))"
done
