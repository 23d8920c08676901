cd /root/repo
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_factorize|k_factor_norms|k_solve_dense" --launch-skip 3 -c 3 -o gpurun_out/final_r1q python bench.py --steps 1 --warmup 1 --nmat 1036 --no-cpu-baseline > gpurun_out/ncu_full_q.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_full_q.log | cut -c1-200
