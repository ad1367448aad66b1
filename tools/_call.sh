timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py 128 gpu 240 3 2>&1 | tail -2
