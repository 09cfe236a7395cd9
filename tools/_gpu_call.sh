mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv2_gpu.py tests/test_mc_gpu.py -m gpu -q --timeout 300 2>&1 | grep -v Warning | grep "^E \|passed\|failed\|FAILED" | head -30 > gpurun_out/r2_t8.log; cat gpurun_out/r2_t8.log
timeout 900 python tools/conv2_microbench.py --out gpurun_out/r2_conv2_mb4.json > gpurun_out/r2_conv2_mb4.txt 2>&1; head -64 gpurun_out/r2_conv2_mb4.txt; tail -1 gpurun_out/r2_conv2_mb4.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-reference --no-fine-tune-api 2>gpurun_out/r2_bench_err.log | tail -1 | tee gpurun_out/r2_bench7.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['final_loss'], d['gpu_launches'])"
