cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; wc -l gpurun_out/bench_n2.json; cut -c1-200 gpurun_out/bench_n2.json
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; wc -l gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_n1.json
