python -m pytest tests/test_gpu_pixel_ops.py tests/test_gpu_shim.py tests/test_gpu_raster.py tests/test_gpu_draw.py -x -q -m gpu 2>&1 | tail -8
