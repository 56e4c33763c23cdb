python -m pytest tests -x -q -m gpu 2>&1 | tail -2
DFPSR_ASYNC=1 python tools/tile_ab.py 256 --tiny 2>&1 | grep -E "exact|tiny"
