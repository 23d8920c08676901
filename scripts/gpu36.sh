cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python scripts/structure_hunt.py 20000 300 700 2>&1 | tail -12 | tee gpurun_out/structure_hunt.txt
timeout 600 python scripts/tunables_hunt.py 7000 200 1500 2>&1 | tail -5 | tee gpurun_out/tunables_hunt2.txt
timeout 600 python bench.py > gpurun_out/bench_r1v.json 2> gpurun_out/bench_r1v.err; echo rc=$?; cut -c1-400 gpurun_out/bench_r1v.json
