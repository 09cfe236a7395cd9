mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python -m pytest tests/test_mc_gpu.py -m gpu -q --timeout 500 -k "bench" 2>&1 | grep -v Warning | grep "^E \|224x384\|graph vs\|passed\|failed" | head -40
done > gpurun_out/r2_mc_t3.log 2>&1
CVD_CONV2=0 timeout 900 python -m pytest tests/test_mc_gpu.py -m gpu -q --timeout 500 -k "bench" 2>&1 | grep -v Warning | grep "^E \|224x384\|graph vs\|passed\|failed" | head -40 >> gpurun_out/r2_mc_t3.log
cat gpurun_out/r2_mc_t3.log
