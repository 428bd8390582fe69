# round-2 GPU batch 10 (2 GPUs): data-parallel tests on the real kernels, then 2-GPU bench lines with the three-part overlapped all-reduce
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_path.py -q -m gpu -x -k "two_rank or graphed or fused" > gpurun_out/r2k_pytest.log 2>&1; tail -5 gpurun_out/r2k_pytest.log
bash scripts/r2_gpu_scale.sh 2 r2k c2 c4
