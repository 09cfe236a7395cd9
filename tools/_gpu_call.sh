mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv2_gpu.py -m gpu -q --timeout 300 2>&1 | grep -v Warning | tail -30 > gpurun_out/r2_conv2_t2.log; tail -30 gpurun_out/r2_conv2_t2.log
timeout 900 python tools/debug_v2_race.py > gpurun_out/r2_race1.log 2>&1; tail -30 gpurun_out/r2_race1.log
