cd /root/repo
rm -f gpurun_out/configs_c5full.jsonl
timeout 1300 python scripts/configs_bench.py --out gpurun_out/configs_c5full.jsonl c5 > gpurun_out/cfg_c5full.log 2>&1; echo "rc $?"
head -1 gpurun_out/configs_c5full.jsonl | cut -c1-1200
