cd /root/repo
for w in 200000 300000 500000 900000; do W_MEM=$w timeout 300 python tests/gpu_norms_timing.py 2>&1 | grep "factorize total"; done
L_MEM=60000 U_MEM=60000 W_MEM=500000 timeout 300 python tests/gpu_norms_timing.py 2>&1 | grep "factorize total"
