set -x
cd /root/repo
mkdir -p gpurun_out
for v in base r4 b6 b8 r4b8 r4b12; do
  if [ $v = base ]; then unset BLU_B200_LIB; else export BLU_B200_LIB=/root/repo/variants/$v.so; fi
  echo "=== variant $v"
  timeout 300 python tests/gpu_timing.py 2048 64,128,256 2>&1 | grep -E "^nt=|phases"
done > gpurun_out/sweep_r1b.log 2>&1
cat gpurun_out/sweep_r1b.log
