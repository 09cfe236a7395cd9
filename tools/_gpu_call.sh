mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv2_gpu.py tests/test_mc_gpu.py -m gpu -q --timeout 300 2>&1 | grep -v Warning | grep "^E \|passed\|failed\|FAILED" | head -30 > gpurun_out/r2_t6.log; cat gpurun_out/r2_t6.log
timeout 900 python tools/conv2_microbench.py --out gpurun_out/r2_conv2_mb3.json > gpurun_out/r2_conv2_mb3.txt 2>&1; grep "k 1\|totals" gpurun_out/r2_conv2_mb3.txt | head -40
for e in "CVD_BNBWD_BLOCKS=3" "CVD_BNBWD_BLOCKS=6"; do
env $e timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-reference 2>gpurun_out/r2_bench_err.log | tail -1 | tee gpurun_out/r2_bench5.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', d['value'], d['ms_per_step'], d['e2e'], d['final_loss'], d['gpu_launches'])"
done
timeout 300 python tools/profile_engine.py --workload mc --out gpurun_out/r2_mc_ops_v5.json 2>&1 | tail -16
