cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/configs_q3.jsonl
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for c in c4 c5; do timeout 900 python scripts/configs_bench.py --quick --out gpurun_out/configs_q3.jsonl $c > gpurun_out/cfgq3_$c.log 2>&1; echo "rc $c $?"; done
cat gpurun_out/configs_q3.jsonl | cut -c1-1800
