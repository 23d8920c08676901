cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python scripts/structure_hunt.py 70000 1200 1500 --nofork --nupd=200 2>&1 | tail -8 | tee gpurun_out/structure_hunt_replay2.txt
timeout 300 python scripts/tunables_hunt.py 80000 600 2500 2>&1 | tail -5 | tee gpurun_out/tunables_hunt3.txt
