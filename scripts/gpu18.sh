cd /root/repo
for v in prev nomerge base; do
  if [ $v = base ]; then unset BLU_B200_LIB; else export BLU_B200_LIB=/root/repo/variants/$v.so; fi
  echo "=== $v"; timeout 300 python tests/gpu_norms_timing.py 2>&1 | tail -2 | head -1
done
