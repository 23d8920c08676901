set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_factorize --launch-skip 1 -c 1 -o gpurun_out/factorize_r1f python bench.py --steps 1 --warmup 1 --nmat 1184 --no-cpu-baseline > gpurun_out/ncu_full_f.log 2>&1; echo rc=$?; tail -3 gpurun_out/ncu_full_f.log
