mkdir -p gpurun_out
timeout 200 python tools/debug_flowmask.py 2>&1 | tail -30 | tee gpurun_out/r2_flowmask_dbg.log
timeout 200 python -m pytest tests/test_flowmask_gpu.py tests/test_consistency_gpu.py -m gpu -q 2>&1 | grep "^E \|passed\|failed" | head
