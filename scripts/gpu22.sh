cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_update --launch-skip 3 -c 1 -o gpurun_out/update_r1c python scripts/configs_bench.py --quick --out /tmp/x.jsonl c5 > gpurun_out/ncu_update.log 2>&1; echo rc=$?
