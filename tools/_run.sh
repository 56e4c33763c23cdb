python -m pytest tests/test_gpu_pixel_ops.py tests/test_gpu_shim.py -x -q -m gpu 2>&1 | tail -3
python tools/resize_time.py
for v in r8b6 r32b6 r16b4 r16b8 r32b8; do DFPSR_LIB=dfpsr_b200/variants/libdfpsr_b200_$v.so python tools/resize_time.py; done
for sr in 4 6 8 12; do echo small_rows=$sr; DFPSR_SMALL_ROWS=$sr python tools/tile_ab.py 256 2>&1 | grep "batch of" ; done
