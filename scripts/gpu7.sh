set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
(timeout 300 python tests/gpu_timing.py 148 1024,512; timeout 300 python tests/gpu_timing.py 296 512,1024,256; timeout 300 python tests/gpu_timing.py 592 256,512; timeout 300 python tests/gpu_timing.py 1184 128,256) 2>&1 | grep -E "^nt=|phases|pivots" > gpurun_out/sweep_r1e.log; cat gpurun_out/sweep_r1e.log
