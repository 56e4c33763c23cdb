for i in 1 2 3 4 5; do python -m pytest tests -x -q -m gpu 2>&1 | tail -1; done
for i in 1 2 3; do DFPSR_ASYNC=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -1; done
python bench.py --steps 60 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['parity_ok'], d['e2e']['value'], d['clocks'])"
