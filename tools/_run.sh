set -x
python -m pytest tests/test_gpu_raster.py tests/test_gpu_sprite_world.py tests/test_gpu_shim.py -x -q 2>&1 | tail -4
python -m pytest tests/test_gpu_tolerance.py -q -s 2>&1 | grep -E "coverage mismatches|passed|failed" > gpurun_out/tol2.txt
rm -f gpurun_out/sweep2.txt
for v in "" d7 d9 d10 d12; do
  echo "variant=$v" >> gpurun_out/sweep2.txt
  if [ -n "$v" ]; then export DFPSR_LIB=$PWD/dfpsr_b200/variants/libdfpsr_b200_$v.so; fi
  python tools/tile_ab.py 256 2>&1 | grep -E "batch of|single" >> gpurun_out/sweep2.txt
done
unset DFPSR_LIB
python tools/tile_ab.py 256 --tiny 2>&1 | grep tiny >> gpurun_out/sweep2.txt
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 6 -o gpurun_out/r2_tile_v2 -f python tools/tile_ab.py 256 > gpurun_out/ncu_v2.log 2>&1
