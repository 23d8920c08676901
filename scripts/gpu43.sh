cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for i in 1 2 3 4; do timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "test_batch_update_replay or small_memory or pipelined" 2>&1 | tail -1; done
