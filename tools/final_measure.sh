# One-GPU measurement pass of a round: tests, bench line, ncu launch list, ncu --set full captures, microbenches.
# Run on the B200 box through gpurun from the repo root; everything lands in gpurun_out/ (copy what is judged to profiles/).
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -v Warning | tail -12 > gpurun_out/${R}_pytest_gpu.log; tail -4 gpurun_out/${R}_pytest_gpu.log
timeout 1200 python bench.py --steps 30 --warmup 5 2>gpurun_out/${R}_bench_err.log | tail -1 > gpurun_out/${R}_bench_final.json; cut -c1-400 gpurun_out/${R}_bench_final.json
CVD_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench_ncu.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-gpu-reference --no-fine-tune-api > gpurun_out/${R}_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${R}_launches_bench_ncu.csv --steps 2 > gpurun_out/${R}_launch_summary.json; head -30 gpurun_out/${R}_launch_summary.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 3 -c 1 -o gpurun_out/${R}_ncu_conv2_fwd_k11 python tools/conv2_microbench.py --one fwd,64,16,11,224,384 --reps 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad2_kernel -s 3 -c 1 -o gpurun_out/${R}_ncu_wgrad2_k11 python tools/conv2_microbench.py --one wgrad,64,16,11,224,384 --reps 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:consistency_kernel -s 2 -c 1 -o gpurun_out/${R}_ncu_loss_grad python tools/loss_microbench.py --sizes 1080x1920 --batches 16 --reps 3 --warmup 2 > /dev/null 2>&1
timeout 900 python tools/loss_microbench.py --batches 1,2,4,8,16,32,64 --reps 20 --out gpurun_out/${R}_loss_microbench.json > gpurun_out/${R}_loss_microbench.log 2>&1; tail -3 gpurun_out/${R}_loss_microbench.log | cut -c1-200
timeout 600 python tools/conv2_microbench.py --out gpurun_out/${R}_conv2_microbench.json > gpurun_out/${R}_conv2_microbench.txt 2>&1; tail -1 gpurun_out/${R}_conv2_microbench.txt
timeout 600 python tools/conv2_microbench.py --wgrad --out gpurun_out/${R}_wgrad2_microbench.json > gpurun_out/${R}_wgrad2_microbench.txt 2>&1; tail -1 gpurun_out/${R}_wgrad2_microbench.txt
timeout 300 python tools/profile_engine.py --workload mc --out gpurun_out/${R}_mc_ops_events.json 2>&1 | tail -3
timeout 400 python bench.py --workload monodepth2 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/${R}_bench_monodepth2.json; cut -c1-200 gpurun_out/${R}_bench_monodepth2.json
timeout 400 python bench.py --workload midas2 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_midas2.json; cut -c1-200 gpurun_out/${R}_bench_midas2.json
# MiDaS-v2: launch list of one step, the dominant shapes (microbench) and one ncu --set full capture of the wide 1x1 weight gradient
CVD_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_midas_launches_ncu.csv \
    python bench.py --workload midas2 --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-gpu-reference --no-fine-tune-api > gpurun_out/${R}_midas_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${R}_midas_launches_ncu.csv --steps 1 > gpurun_out/${R}_midas_launch_summary.json; head -12 gpurun_out/${R}_midas_launch_summary.json
timeout 300 python tools/midas_microbench.py --out gpurun_out/${R}_midas_microbench.json > gpurun_out/${R}_midas_microbench.txt 2>&1; tail -2 gpurun_out/${R}_midas_microbench.txt | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 -o gpurun_out/${R}_ncu_midas_wgrad_1x1 -f python tools/midas_microbench.py --only 1024,1024,1,24,42 --reps 2 > /dev/null 2>&1
