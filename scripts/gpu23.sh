cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python tests/gpu_update_timing.py 2>&1 | tail -5
