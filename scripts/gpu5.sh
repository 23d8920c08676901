set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python tests/gpu_timing.py 2048 32,64,128 2>&1 | grep -E "^nt=|phases|pivots" > gpurun_out/sweep_r1c.log; cat gpurun_out/sweep_r1c.log
python - <<'PY'
import sys; sys.path.insert(0,'.')
from blu_b200 import BLUBatch, gen
import numpy as np
nmat,m=64,2000
bb,be,bi,bx,rhs=gen.batch(nmat,m,700,5.0,2000,3000)
b=BLUBatch(nmat,m,int((be-bb).reshape(nmat,m).sum(1).max()))
b.l_mem=100000;b.u_mem=100000;b.w_mem=160000
b.upload(bb,be,bi,bx,rhs); b.factorize_resident()
print("final mem l,u,w", b.get_param("l_mem"), b.get_param("u_mem"), b.get_param("w_mem"), "nrealloc", b.info(0,"nrealloc"))
print("max l_nz,u_nz", max(b.info(k,"l_nz") for k in range(nmat)), max(b.info(k,"u_nz") for k in range(nmat)), "ngarbage", [b.info(k,"ngarbage") for k in range(8)], "nexpand", [b.info(k,"nexpand") for k in range(4)])
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_factorize --launch-skip 3 -c 1 -o gpurun_out/factorize_r1c python bench.py --steps 1 --warmup 1 --nmat 1184 --no-cpu-baseline > gpurun_out/ncu_full_c.log 2>&1; echo rc=$?; tail -3 gpurun_out/ncu_full_c.log
