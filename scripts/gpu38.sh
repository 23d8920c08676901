cd /root/repo
mkdir -p gpurun_out
timeout 500 python scripts/structure_hunt.py 60000 240 500 --nofork --nupd=150 2>&1 | tail -8 | tee gpurun_out/structure_hunt_replay.txt
