set -x
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; echo rc=$?; tail -3 gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1a.json 2>&1; cat gpurun_out/bench_ref_r1a.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_list.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_factorize -c 1 -o gpurun_out/factorize_r1a python bench.py --steps 1 --warmup 1 --nmat 592 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
