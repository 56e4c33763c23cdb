set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
DFPSR_ASYNC=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for chain in 0 1; do for as in 0 1; do echo "chain=$chain async=$as"; DFPSR_CHAIN=$chain DFPSR_ASYNC=$as python tools/single_frame_profile.py 2>&1 | grep wall; done; done
DFPSR_ASYNC=1 DFPSR_SW_TIMING=1 DFPSR_END_TIMING=1 python tools/sprite_world_profile.py 2>&1 | tail -8
DFPSR_ASYNC=1 python tools/tile_ab.py 256 --tiny 2>&1 | grep -E "batch of|single|tiny"
