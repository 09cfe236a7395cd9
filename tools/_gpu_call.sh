mkdir -p gpurun_out
timeout 900 python tools/conv2_microbench.py --wgrad --out gpurun_out/r2_wgrad2_mb1.json > gpurun_out/r2_wgrad2_mb1.txt 2>&1; tail -45 gpurun_out/r2_wgrad2_mb1.txt
timeout 900 python -m pytest tests/test_mc_gpu.py tests/test_parameter_loss_gpu.py tests/test_fine_tuner_2rank_gpu.py tests/test_fine_tuner_gpu.py -m gpu -q --timeout 500 2>&1 | grep -v Warning | grep "^E \|224x384\|graph vs\|passed\|failed\|FAILED" | head -40 > gpurun_out/r2_mc_t4.log; cat gpurun_out/r2_mc_t4.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-reference 2>gpurun_out/r2_bench_err.log | tail -1 | tee gpurun_out/r2_bench3.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['final_loss'], d['gpu_launches'])"
timeout 300 python tools/profile_engine.py --workload mc --out gpurun_out/r2_mc_ops_v3.json 2>&1 | tail -20
