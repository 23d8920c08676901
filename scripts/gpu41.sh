cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "test_batch_update_replay and True" 2>&1 | grep -E "^E |passed|failed|Error" | head -12; done > gpurun_out/flaky.txt 2>&1
cat gpurun_out/flaky.txt | cut -c1-400
