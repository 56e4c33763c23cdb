python -m pytest tests/test_gpu_pixel_ops.py tests/test_gpu_raster.py -x -q -m gpu -k "strided or wireframe" 2>&1 | tail -15
