mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -v Warning | tail -15 > gpurun_out/r2_pytest_full1.log; tail -8 gpurun_out/r2_pytest_full1.log
timeout 300 python tools/profile_engine.py --workload mc --out gpurun_out/r2_mc_ops_v7.json 2>&1 | tail -16
timeout 900 python bench.py --steps 30 --warmup 5 2>gpurun_out/r2_bench_full_err.log | tail -1 > gpurun_out/r2_bench_full1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_full1.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['final_loss'])
print('roofline', {k:d['roofline'][k] for k in ('kernel','achieved','frac','launches')}, d['roofline']['all_conv_kernels'])
for k,v in d['roofline']['by_kernel'].items(): print('  ',k[:30], v)
print('loss roof', d['roofline_loss_kernel'])
print('cpu', d['cpu_baseline']); print('gpu_ref', d['gpu_reference']); print('api', d['fine_tune_api']); print('parity', d['parity_after_steps']); print(d['clocks'])
"; tail -5 gpurun_out/r2_bench_full_err.log
