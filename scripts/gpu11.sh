set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/configs_r1.jsonl
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; echo rc=$?; tail -3 gpurun_out/bench_r1j.err; cat gpurun_out/bench_r1j.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1j.json 2>&1; cat gpurun_out/bench_ref_r1j.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_j.log 2>&1; echo rc=$?
for c in c1 c4 c3 c5; do
  timeout 1200 python scripts/configs_bench.py --out gpurun_out/configs_r1.jsonl $c > gpurun_out/cfgfull_$c.log 2>&1; echo "rc $c $?"; tail -c 1500 gpurun_out/cfgfull_$c.log
done
