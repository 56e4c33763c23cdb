python -m pytest tests/test_gpu_sprite_world.py -x -q -m gpu 2>&1 | tail -15
