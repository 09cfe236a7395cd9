#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_full_tests.log 2>&1
tail -3 gpurun_out/r2_full_tests.log
for wl in midas2 monodepth2 mc; do
  python bench.py --workload $wl --steps 20 --warmup 5 --no-gpu-reference --no-fine-tune-api --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_pack_$wl.json
  python -c "
import json; j = json.load(open('gpurun_out/r2_pack_$wl.json')); print('$wl', j['value'], j['ms_per_step'], j['gpu_launches'])"
done
python tools/profile_engine.py --workload midas2 --out gpurun_out/r2_midas_ops3.json 2>&1 | tail -1
