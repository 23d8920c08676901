set -x
cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo rc=$?; tail -3 gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
