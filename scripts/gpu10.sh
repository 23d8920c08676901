cd /root/repo
mkdir -p gpurun_out
run() { echo "=== $1 cap=$2"; if [ $1 = base ]; then unset BLU_B200_LIB; else export BLU_B200_LIB=/root/repo/variants/$1.so; fi; BLU_B200_CAP=$2 timeout 300 python tests/gpu_timing.py 2048 $3 2>&1 | grep -E "^nt="; }
(run b7 256 128; run b7r6 256 128; run b7r7 256 128; run b8r6 256 128; run b7r6 192 128; run b7 256 96; run b7 256 160) > gpurun_out/sweep_r1i.log 2>&1
cat gpurun_out/sweep_r1i.log
