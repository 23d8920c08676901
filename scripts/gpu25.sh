cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -12
