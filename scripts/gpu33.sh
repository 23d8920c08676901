cd /root/repo
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/final2_bench.json)"; cut -c1-260 gpurun_out/final2_bench.json
