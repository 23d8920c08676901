cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$? lines=$(wc -l < gpurun_out/final_ref.json)"; cut -c1-220 gpurun_out/final_ref.json
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/final_bench.json)"; cat gpurun_out/final_bench.json
