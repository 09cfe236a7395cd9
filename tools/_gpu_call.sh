#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_gpu.py tests/test_midas_gpu.py tests/test_mono2_gpu.py -x -q -m gpu > gpurun_out/r2_chunk_tests.log 2>&1
tail -3 gpurun_out/r2_chunk_tests.log
L=gpurun_out/r2_midas_mb2.log; : > $L
echo "== v4" >> $L; python tools/midas_microbench.py >> $L 2>&1
echo "== scalar" >> $L; CVD_WG_NO_V4=1 python tools/midas_microbench.py >> $L 2>&1
python bench.py --workload midas2 --steps 10 --warmup 3 --no-gpu-reference --no-fine-tune-api 2>/dev/null | tail -1 > gpurun_out/r2_midas_b.json
python -c "
import json; j = json.load(open('gpurun_out/r2_midas_b.json')); print('midas', j['value'], j['ms_per_step'], j['gpu_launches'])"
python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-fine-tune-api --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_chunk_mc.json
python -c "
import json; j = json.load(open('gpurun_out/r2_chunk_mc.json')); print('mc', j['value'], j['ms_per_step'])"
