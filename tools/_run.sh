SEL='not golden and not 1080 and not 4k and not tiny and not large and not 8192 and not rsqrt and not returns_before and not pools_held'
FILES="tests/test_gpu_sprite_world.py tests/test_gpu_draw.py tests/test_gpu_raster.py tests/test_gpu_pixel_ops.py tests/test_gpu_async.py tests/test_gpu_tolerance.py"
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $FILES -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|SUMMARY" | tail -3
done
echo "== memcheck async"; DFPSR_ASYNC=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sprite_world.py tests/test_gpu_raster.py tests/test_gpu_async.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|SUMMARY" | tail -3
