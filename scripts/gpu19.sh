cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/configs_q2.jsonl
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for c in c1 c3; do timeout 600 python scripts/configs_bench.py --quick --out gpurun_out/configs_q2.jsonl $c > gpurun_out/cfgq2_$c.log 2>&1; echo "rc $c $?"; done
python - <<'PY'
import json
for l in open('gpurun_out/configs_q2.jsonl'):
    r=json.loads(l); print(r["config"][:60], "| GPU fact %.1f ms solve %.2f ms | CPU fact %.1f ms solve %.2f ms | parity %s" % (r["gpu"]["factorize_ms_wall"], r["gpu"]["solve_dense_ms_wall"], r["cpu_oracle_1thread"]["factorize_ms"], r["cpu_oracle_1thread"]["solve_dense_ms"], r["parity"]))
PY
timeout 300 python tests/gpu_norms_timing.py 2>&1 | tail -2 | head -1
