set -x
mkdir -p gpurun_out
timeout 120 python profiles/diag/loopback_probe.py 2 5 > gpurun_out/r2c_probe_default.log 2>&1
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 python profiles/diag/loopback_probe.py 2 5 > gpurun_out/r2c_probe_conn32.log 2>&1
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 python profiles/diag/loopback_probe.py 8 5 > gpurun_out/r2c_probe_conn32_w8.log 2>&1
cat gpurun_out/r2c_probe_*.log
