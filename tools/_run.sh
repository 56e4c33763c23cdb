python tools/tiny_profile.py
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
DFPSR_ASYNC=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
