set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python tests/gpu_timing.py 2>&1 | tail -40
