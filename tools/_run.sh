set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
DFPSR_ASYNC=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python -m pytest tests/test_gpu_async.py -q -s 2>&1 | grep -E "returned after|passed|failed"
python tools/tile_ab.py 256 --tiny 2>&1 | grep -E "batch of|single|tiny" > gpurun_out/sweep3.txt
echo ASYNC >> gpurun_out/sweep3.txt
DFPSR_ASYNC=1 python tools/tile_ab.py 256 --tiny 2>&1 | grep -E "batch of|single|tiny" >> gpurun_out/sweep3.txt
cat gpurun_out/sweep3.txt
