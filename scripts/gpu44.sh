cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done | tee gpurun_out/suite_repeats.txt
