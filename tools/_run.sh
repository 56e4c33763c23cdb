python -m pytest tests -x -q -m gpu 2>&1 | tail -3
DFPSR_ASYNC=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
DFPSR_ASYNC=1 python tools/sprite_world_profile.py 2>&1 | tail -2
DFPSR_ASYNC=1 python tools/tile_ab.py 256 2>&1 | grep -E "single"
python - <<'PY'
import sys, os, json
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "tests")]
from dfpsr_b200 import lib
cuda = lib.load(); lib.check(cuda.dfpsr_init(0)); lib.check(cuda.dfpsr_set_default_async(1))
import bench_extras
out = bench_extras.run(cuda, lib, cpu=True)
print(json.dumps({k: out[k] for k in ("sandbox_800x600_sprite_world", "terrain_1080p_single_frame", "tiny_triangles_4k")}, indent=1))
PY
