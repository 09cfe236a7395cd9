set -x
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -40 > gpurun_out/r2_pytest1.log; tail -15 gpurun_out/r2_pytest1.log
for e in "" "CVD_FORK_FIRST=1" "CVD_KXFWD=1"; do
  env $e timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-gpu-reference 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', d['value'], d['ms_per_step'], d['e2e'])"
done 2>&1 | tee gpurun_out/r2_ab1.log
timeout 600 python bench.py --impl reference-gpu --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_refgpu1.json
