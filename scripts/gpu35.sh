cd /root/repo
mkdir -p gpurun_out
timeout 1200 python scripts/tunables_hunt.py 5000 400 1500 2>&1 | tail -15 | tee gpurun_out/tunables_hunt.txt
