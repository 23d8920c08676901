cd /root/repo
rm -f gpurun_out/configs_q4.jsonl
timeout 600 python scripts/configs_bench.py --quick --out gpurun_out/configs_q4.jsonl c5 > gpurun_out/cfgq4_c5.log 2>&1; echo "rc $?"
head -1 gpurun_out/configs_q4.jsonl | cut -c1-900
