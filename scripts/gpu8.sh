set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/configs_quick.jsonl
timeout 120 python tests/gpu_timing.py 148 512 2>&1 | grep -i -E "error|nt=" | head -5
for c in c1 c3 c4 c5; do
  timeout 600 python scripts/configs_bench.py --quick --out gpurun_out/configs_quick.jsonl $c > gpurun_out/cfg_$c.log 2>&1; echo "rc $c $?"; tail -c 2500 gpurun_out/cfg_$c.log
done
