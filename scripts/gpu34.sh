cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1t.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_t.log 2>&1; echo rc=$?
