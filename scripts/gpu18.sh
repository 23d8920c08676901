cd /root/repo
for v in base pfl1 base pfl1; do
  if [ $v = base ]; then unset BLU_B200_LIB; else export BLU_B200_LIB=/root/repo/variants/$v.so; fi
  echo "=== $v"; timeout 300 python tests/gpu_norms_timing.py 2>&1 | tail -2 | head -1
done
