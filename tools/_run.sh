python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; head -c 400 gpurun_out/bench_quick.json; echo
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 600
