set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --n 256 --steps 5 --no-cpu 2>&1 | tail -3
python bench.py --steps 10 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2
