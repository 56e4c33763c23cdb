python tools/tiny_profile.py
python tools/tiny_profile.py
