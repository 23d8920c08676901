set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo rc=$?; tail -3 gpurun_out/bench_r1m.err; cat gpurun_out/bench_r1m.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1m.json 2>&1; cat gpurun_out/bench_ref_r1m.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_m.log 2>&1; echo rc=$?
timeout 300 python scripts/configs_bench.py --out gpurun_out/configs_r1m.jsonl c1 > gpurun_out/cfg_m_c1.log 2>&1; tail -c 1200 gpurun_out/cfg_m_c1.log
