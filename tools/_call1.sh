set -x
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/r2_sysinfo.txt 2>&1
lscpu | head -30 >> gpurun_out/r2_sysinfo.txt 2>&1
ls /sys/devices/system/node/ >> gpurun_out/r2_sysinfo.txt 2>&1
free -g >> gpurun_out/r2_sysinfo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 300 python tools/copy_ceiling_mp.py --gpus 1 > gpurun_out/r2_ceiling_n1.txt 2>&1
tail -5 gpurun_out/r2_pytest1.log; cat gpurun_out/r2_ceiling_n1.txt
