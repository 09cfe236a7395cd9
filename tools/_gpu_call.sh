mkdir -p gpurun_out
run() { env $1 timeout 120 python tools/conv2_microbench.py --one $2 --reps 5 2>&1 | tail -1 | sed "s/^/[$1] /"; }
for one in fwd,64,16,11,224,384 fwd,32,32,7,112,192 fwd,64,32,7,112,192; do
  for e in "X=0" "CVD2_NB=3" "CVD2_NB=4" "CVD2_ASPLIT=1" "CVD2_MT=2" "CVD2_MT=2 CVD2_ASPLIT=1"; do run "$e" $one; done
done 2>&1 | tee gpurun_out/r2_knobs1.log
for one in wgrad,64,16,7,224,384 wgrad,64,16,11,224,384 wgrad,64,32,7,112,192 wgrad,32,32,5,112,192; do
  for e in "CVD2_WG_STAGES=2" "CVD2_WG_STAGES=3" "CVD2_WG_STAGES=4" "CVD2_WG_GPP=1" "CVD2_WG_GPP=1 CVD2_WG_STAGES=3"; do run "$e" $one; done
done 2>&1 | tee -a gpurun_out/r2_knobs1.log
timeout 300 python -m pytest tests/test_consistency_gpu.py -m gpu -q 2>&1 | tail -3
for e in "CVD_LOSS_X4=0" "CVD_LOSS_X4=1"; do env $e timeout 200 python tools/loss_microbench.py --sizes 224x384,1080x1920 --batches 4,16 --reps 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$e', d['H'], d['B'], d['mode'], round(d['ms'],4), round(d['GBps']), round(d['frac_of_peak'],3))
"; done 2>&1 | tee gpurun_out/r2_loss1.log
