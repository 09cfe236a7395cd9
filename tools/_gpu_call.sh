mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv2_gpu.py tests/test_mc_gpu.py -m gpu -q --timeout 300 2>&1 | grep -v Warning | grep "^E \|passed\|failed\|FAILED" | head -30 > gpurun_out/r2_t9.log; cat gpurun_out/r2_t9.log
timeout 900 python tools/conv2_microbench.py --out gpurun_out/r2_conv2_mb5.json > gpurun_out/r2_conv2_mb5.txt 2>&1; head -34 gpurun_out/r2_conv2_mb5.txt | cut -c1-110; tail -1 gpurun_out/r2_conv2_mb5.txt
timeout 900 python tools/conv2_microbench.py --wgrad --out gpurun_out/r2_wgrad2_mb2.json > gpurun_out/r2_wgrad2_mb2.txt 2>&1; cat gpurun_out/r2_wgrad2_mb2.txt | cut -c1-110
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-reference --no-fine-tune-api 2>gpurun_out/r2_bench_err.log | tail -1 | tee gpurun_out/r2_bench8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['final_loss'], d['gpu_launches'])"
