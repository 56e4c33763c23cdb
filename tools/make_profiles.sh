#!/bin/bash
# Round profile artefacts (run on a GPU box through gpurun; outputs under gpurun_out/, summarised into profiles/ by tools/summarise_profiles.py):
#   1. the headline bench line                                   gpurun_out/bench_final.json
#   2. ncu launch list of the same command                       gpurun_out/launches.csv
#   3. ncu --set full of the tile kernel's launch, both modes    gpurun_out/tile_exact.ncu-rep, tile_tolerance.ncu-rep (256 views in one launch)
#   4. ncu --set full of the set-up side of the same launch      gpurun_out/setup_batch.ncu-rep
set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 2 -c 1 -o gpurun_out/tile_exact -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_exact.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 2 -c 1 -o gpurun_out/tile_tolerance -f python bench.py --precision tolerance --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_tolerance.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"big_units|setup_kernel|counts_kernel" -s 8 -c 4 -o gpurun_out/setup_batch -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_setup.log 2>&1
cp dfpsr_b200/csrc/raster.cu gpurun_out/raster_profiled.cu
ls -la gpurun_out/
