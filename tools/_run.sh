python -m pytest tests/test_gpu_raster.py tests/test_gpu_sprite_world.py tests/test_gpu_shim.py tests/test_gpu_async.py -x -q -m gpu 2>&1 | tail -3
python tools/tile_ab.py 256 --tiny 2>&1 | grep "exact"
