mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flowmask_gpu.py tests/test_mc_gpu.py tests/test_fine_tuner_gpu.py tests/test_fine_tuner_2rank_gpu.py tests/test_parameter_loss_gpu.py -m gpu -q --timeout 500 2>&1 | grep -v Warning | grep "^E \|passed\|failed\|FAILED\|graph vs" | head -30 > gpurun_out/r2_t10.log; cat gpurun_out/r2_t10.log
for e in "CVD_WGRAD_ASYNC=0" "CVD_WGRAD_ASYNC=1"; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-reference --no-fine-tune-api 2>gpurun_out/r2_bench_err.log | tail -1 | tee gpurun_out/r2_bench9.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', d['value'], d['ms_per_step'], d['e2e'], d['final_loss'], d['gpu_launches'])"
done
