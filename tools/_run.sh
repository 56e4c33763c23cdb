for i in 1 2 3 4 5 6; do python -m pytest tests/test_gpu_raster.py tests/test_gpu_async.py -x -q -m gpu 2>&1 | tail -1; done
for i in 1 2 3 4; do DFPSR_ASYNC=1 python -m pytest tests/test_gpu_raster.py tests/test_gpu_shim.py -x -q -m gpu 2>&1 | tail -1; done
for i in 1 2; do DFPSR_CHAIN=0 python -m pytest tests/test_gpu_raster.py -x -q -m gpu 2>&1 | tail -1; done
