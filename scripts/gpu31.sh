cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python tests/gpu_batch_replay_timing.py 1024 2>&1 | tail -3
timeout 600 python tests/gpu_batch_replay_timing.py 4096 2>&1 | tail -2
