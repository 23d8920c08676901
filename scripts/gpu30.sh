cd /root/repo
rm -f gpurun_out/configs_q7.jsonl
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 900 python scripts/configs_bench.py --quick --nrhs 1000 --out gpurun_out/configs_q7.jsonl c4 > gpurun_out/cfgq7_c4.log 2>&1; echo "rc $?"; tail -3 gpurun_out/cfgq7_c4.log | cut -c1-300
