# round-2 GPU batch 6: regulariser kernels / full stage-2 loss tests, device rays + sharded render tests, bench with --full-loss
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "regulariser or full_stage2 or device_rays or rank_shares or without_samples" > gpurun_out/r2f_pytest_new.log 2>&1; tail -25 gpurun_out/r2f_pytest_new.log
python bench.py --full-loss --no-cpu-baseline > gpurun_out/r2f_c2full.json 2> gpurun_out/r2f_c2full.err; tail -5 gpurun_out/r2f_c2full.err
python bench.py --full-loss --workload c4 --no-cpu-baseline > gpurun_out/r2f_c4full.json 2> gpurun_out/r2f_c4full.err; tail -5 gpurun_out/r2f_c4full.err
