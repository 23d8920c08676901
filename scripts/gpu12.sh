set -x
cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; echo rc=$?; tail -3 gpurun_out/bench_r1k.err; cat gpurun_out/bench_r1k.json
timeout 120 python tests/gpu_timing.py 296 512 2>&1 | grep -E "^nt=|rror"
