python -m pytest tests/test_gpu_raster.py -x -q -m gpu 2>&1 | tail -2
python tools/tile_ab.py 256 2>&1 | grep batch
echo "== 16x8 tiles"
DFPSR_LIB=dfpsr_b200/variants/libdfpsr_b200_t16.so python -m pytest tests/test_gpu_raster.py -q -m gpu 2>&1 | tail -8
DFPSR_LIB=dfpsr_b200/variants/libdfpsr_b200_t16.so python tools/tile_ab.py 256 --tiny 2>&1
