mkdir -p gpurun_out
for b in 4 8 16; do
CVD_MIDAS_BRANCHES=$b timeout 300 python bench.py --workload midas2 --steps 8 --warmup 3 --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('branches $b', d['value'], d['ms_per_step'])"
done 2>&1 | tee gpurun_out/r2_midas_branches.log
