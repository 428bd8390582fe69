# round-2 GPU batch 11: whole suite; compute-sanitizer memcheck + racecheck of smoke(); ncu --set full of the HBM-bound kernels
# (scripts/kernel_roofline.py: lbs_fwd, grid build, composite_fwd, adam_multi); c2 launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; tail -6 gpurun_out/r2n_pytest.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py smoke > gpurun_out/r2n_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2n_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python __graft_entry__.py smoke > gpurun_out/r2n_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2n_sanitizer_racecheck.log
python scripts/kernel_roofline.py > gpurun_out/r2n_roof.json 2> gpurun_out/r2n_roof.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lbs_fwd_kernel|composite_fwd_kernel|adam_multi_kernel|grid_count_kernel|grid_scatter_kernel|grid_top_kernel" -c 8 -o gpurun_out/r2n_hbm_kernels -f python scripts/kernel_roofline.py > gpurun_out/r2n_ncu_hbm.log 2>&1; tail -2 gpurun_out/r2n_ncu_hbm.log
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --train-path static > gpurun_out/r2n_c2_short.json 2> gpurun_out/r2n_c2_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2n_launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --train-path static > gpurun_out/r2n_ncu_c2.log 2>&1
