#!/bin/bash
# scratch GPU call
mkdir -p gpurun_out
python -m pytest tests/test_midas_gpu.py tests/test_mono2_gpu.py tests/test_conv_gpu.py tests/test_mc_gpu.py -x -q -m gpu > gpurun_out/r2_chunk_tests.log 2>&1
tail -5 gpurun_out/r2_chunk_tests.log
for cl in 1 0; do
  CVD_MIDAS_CHUNK_LAUNCH=$cl python bench.py --workload midas2 --steps 10 --warmup 3 --no-gpu-reference --no-fine-tune-api 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        j = json.loads(l); print('chunk_launch $cl', j['value'], j['ms_per_step'], j.get('gpu_launches'))
" | tee -a gpurun_out/r2_chunk_bench.log
done
python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-fine-tune-api 2>/dev/null | tail -1 > gpurun_out/r2_chunk_mc.json
python -c "
import json; j = json.load(open('gpurun_out/r2_chunk_mc.json')); print('mc', j['value'], j['ms_per_step'])"
python bench.py --workload monodepth2 --steps 10 --warmup 3 --no-gpu-reference --no-fine-tune-api 2>/dev/null | tail -1 | python -c "
import sys, json; j = json.loads(sys.stdin.read()); print('mono2', j['value'], j['ms_per_step'])" | tee -a gpurun_out/r2_chunk_bench.log
