cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_factor_norms --launch-skip 1 -c 1 -o gpurun_out/norms_r1 python tests/gpu_norms_timing.py > gpurun_out/ncu_norms.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_norms.log
