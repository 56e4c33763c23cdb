python -m pytest tests -x -q -m gpu 2>&1 | tail -3
DFPSR_ASYNC=1 python tools/tile_ab.py 256 --tiny 2>&1 | grep -E "batch of|single|tiny"
DFPSR_ASYNC=1 python tools/sprite_world_profile.py 2>&1 | tail -3
