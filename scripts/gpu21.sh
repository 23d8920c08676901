cd /root/repo
rm -f gpurun_out/configs_q4.jsonl
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for c in c5 c4; do timeout 600 python scripts/configs_bench.py --quick --out gpurun_out/configs_q4.jsonl $c > gpurun_out/cfgq4_$c.log 2>&1; echo "rc $c $?"; done
