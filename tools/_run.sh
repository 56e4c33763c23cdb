python -m pytest tests/test_gpu_raster.py -x -q -k "sdk or broad" 2>&1 | tail -2
SEL='not golden and not 1080 and not 4k and not tiny and not large and not 8192 and not rsqrt and not returns_before'
for tool in memcheck racecheck initcheck; do
  echo "== $tool sync"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_sprite_world.py tests/test_gpu_draw.py tests/test_gpu_raster.py tests/test_gpu_pixel_ops.py tests/test_gpu_async.py tests/test_gpu_tolerance.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -4
done
echo "== memcheck async"
DFPSR_ASYNC=1 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sprite_world.py tests/test_gpu_raster.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -3
