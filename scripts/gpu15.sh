cd /root/repo
for v in n9 n12 n14; do
  export BLU_B200_LIB=/root/repo/variants/$v.so
  echo "=== $v"; timeout 300 python tests/gpu_norms_timing.py 2>&1 | tail -2 | head -1
done
