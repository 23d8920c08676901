cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tests/gpu_norms_timing.py 2>&1 | tail -2
