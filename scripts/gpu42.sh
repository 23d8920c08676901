cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "test_batch_update_replay" 2>&1 | grep -E "^E |passed|failed|Error" | head -6; done 2>&1 | tee gpurun_out/flaky_after_fix.txt | cut -c1-300
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; echo rc=$?; cut -c1-300 gpurun_out/bench_r1w.json
