mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv2_gpu.py tests/test_mc_gpu.py -m gpu -q --timeout 300 2>&1 | grep -v Warning | grep "^E \|passed\|failed\|FAILED" | head -30 > gpurun_out/r2_t7.log; cat gpurun_out/r2_t7.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-reference 2>gpurun_out/r2_bench_err.log | tail -1 | tee gpurun_out/r2_bench6.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['final_loss'], d['gpu_launches'])"
timeout 300 python tools/profile_engine.py --workload mc --out gpurun_out/r2_mc_ops_v6.json 2>&1 | tail -9
i=0
for one in fwd,32,32,7,112,192 fwd,64,16,11,224,384 fwd,128,208,1,224,384 wgrad,64,16,7,224,384 wgrad,64,16,11,224,384; do
  i=$((i+1)); kern=conv2_kernel; case $one in wgrad*) kern=wgrad2_kernel;; esac
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -o gpurun_out/r2_ncu_$i python tools/conv2_microbench.py --one $one --reps 2 > gpurun_out/r2_ncu_$i.log 2>&1
  tail -2 gpurun_out/r2_ncu_$i.log
done
ls -la gpurun_out/*.ncu-rep | tail -8
