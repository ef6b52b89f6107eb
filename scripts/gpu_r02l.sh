#!/bin/bash
# r02l: full GPU test suite + default bench line after the row-lane velocity kernel
T=${1:-r02l}
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-solve > gpurun_out/${T}_bench_t3d92.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_t3d92.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["parity_check"])
PY
