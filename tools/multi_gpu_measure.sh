# Multi-GPU measurement pass (run through `gpurun --gpus 8`): weak scaling of the headline workload at 8 GPUs, BASELINE
# configs[2] (MiDaS-v2, global batch 8 on 8 GPUs) and configs[3] (monodepth2, global batch 16 on 4 GPUs), NCCL evidence.
# Second argument: "c3" = only the MiDaS line (gpurun --gpus 8), "c4" = only the monodepth2 line (gpurun --gpus 4), default all.
R=${1:-r02}
ONLY=${2:-all}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
if [ "$ONLY" = "c3" ]; then
  timeout 900 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --workload midas2 --steps 10 --warmup 3 --no-roofline 2>gpurun_out/${R}_midas2_8gpu.err | grep '^{' | tail -1 > gpurun_out/${R}_bench_midas2_8gpu.json; cut -c1-300 gpurun_out/${R}_bench_midas2_8gpu.json
  exit 0
fi
if [ "$ONLY" = "c4" ]; then
  timeout 900 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --workload monodepth2 --steps 20 --warmup 5 --no-roofline 2>gpurun_out/${R}_mono2_4gpu.err | grep '^{' | tail -1 > gpurun_out/${R}_bench_monodepth2_4gpu.json; cut -c1-300 gpurun_out/${R}_bench_monodepth2_4gpu.json
  exit 0
fi
NCCL_DEBUG=INFO timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/${R}_bench_8gpu.log 2>&1
grep '^{' gpurun_out/${R}_bench_8gpu.log | tail -1 > gpurun_out/${R}_bench_8gpu.json; cut -c1-300 gpurun_out/${R}_bench_8gpu.json
grep -E "NCCL INFO (Connected|.*NVLS|comm .* rank|Channel 00/|ncclCommInitRank)" gpurun_out/${R}_bench_8gpu.log | head -40 > gpurun_out/${R}_nccl_evidence_8gpu.txt; wc -l gpurun_out/${R}_nccl_evidence_8gpu.txt
timeout 900 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --workload midas2 --steps 10 --warmup 3 2>gpurun_out/${R}_midas2_8gpu.err | grep '^{' | tail -1 > gpurun_out/${R}_bench_midas2_8gpu.json; cut -c1-300 gpurun_out/${R}_bench_midas2_8gpu.json
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --workload monodepth2 --steps 20 --warmup 5 2>gpurun_out/${R}_mono2_4gpu.err | grep '^{' | tail -1 > gpurun_out/${R}_bench_monodepth2_4gpu.json; cut -c1-300 gpurun_out/${R}_bench_monodepth2_4gpu.json
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_fine_tuner_2rank_gpu.py -m gpu -q 2>&1 | tail -2
