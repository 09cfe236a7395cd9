#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_midas_gpu.py tests/test_mono2_gpu.py -x -q -m gpu > gpurun_out/r2_pair_tests.log 2>&1
tail -3 gpurun_out/r2_pair_tests.log
for pw in 1 0; do for wl in midas2 monodepth2; do
  CVD_PAIR_WG=$pw python bench.py --workload $wl --steps 20 --warmup 5 --no-gpu-reference --no-fine-tune-api --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_pair_${wl}_$pw.json
  python -c "
import json; j = json.load(open('gpurun_out/r2_pair_${wl}_$pw.json')); print('pair $pw $wl', j['value'], j['ms_per_step'], j['gpu_launches'])"
done; done
