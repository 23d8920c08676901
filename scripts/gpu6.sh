set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python tests/gpu_timing.py 2048 64,128,256 2>&1 | grep -E "^nt=|phases|pivots" > gpurun_out/sweep_r1d.log; cat gpurun_out/sweep_r1d.log
