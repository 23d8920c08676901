cd /root/repo
for v in prev base prev base; do
  if [ $v = base ]; then unset BLU_B200_LIB; else export BLU_B200_LIB=/root/repo/variants/$v.so; fi
  echo "=== $v"; W_MEM=900000 timeout 300 python tests/gpu_norms_timing.py 2>&1 | grep "factorize total"
done
