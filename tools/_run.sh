ncu --set full --clock-control none --import-source on -k regex:"setup_kernel|raster_kernel" -s 14 -c 3 -o gpurun_out/r2_tiny -f python tools/tiny_profile.py > gpurun_out/ncu_tiny_r2.log 2>&1
tail -2 gpurun_out/ncu_tiny_r2.log
cp dfpsr_b200/csrc/raster.cu gpurun_out/raster_profiled_tiny.cu
