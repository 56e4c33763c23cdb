python -m pytest tests -x -q -m gpu 2>&1 | tail -3
DFPSR_ASYNC=1 python -m pytest tests/test_gpu_raster.py tests/test_gpu_pixel_ops.py -x -q 2>&1 | tail -2
