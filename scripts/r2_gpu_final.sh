# round-2 final measurement batch (1 GPU): whole suite, the bench lines behind profiles/r02_bench_*, the reference arm, the c2 launch
# list, the per-kernel HBM roofline and ncu --set full of the HBM-bound kernels.  Every command under ncu ran without it first.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; tail -4 gpurun_out/r2z_pytest.log
python bench.py > gpurun_out/r2z_c2.json 2> gpurun_out/r2z_c2.err; tail -2 gpurun_out/r2z_c2.err
python bench.py --full-loss --no-cpu-baseline > gpurun_out/r2z_c2full.json 2> gpurun_out/r2z_c2full.err; tail -2 gpurun_out/r2z_c2full.err
python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2z_c4.json 2> gpurun_out/r2z_c4.err; tail -2 gpurun_out/r2z_c4.err
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r2z_c1.json 2> gpurun_out/r2z_c1.err; tail -2 gpurun_out/r2z_c1.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2z_c3.json 2> gpurun_out/r2z_c3.err; tail -2 gpurun_out/r2z_c3.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_c5.json 2> gpurun_out/r2z_c5.err; tail -2 gpurun_out/r2z_c5.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 --ref-budget 60 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; tail -2 gpurun_out/r2z_ref.err
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --train-path static > gpurun_out/r2z_c2_short.json 2> gpurun_out/r2z_c2_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2z_launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --train-path static > gpurun_out/r2z_ncu_c2.log 2>&1
python scripts/kernel_roofline.py > gpurun_out/r2z_roof.json 2> gpurun_out/r2z_roof.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:composite_fwd_kernel|adam_multi_kernel|grid_count_kernel|grid_scatter_kernel|grid_top_kernel|grid_header_kernel" -c 30 -o gpurun_out/r2z_hbm_kernels -f python scripts/kernel_roofline.py > gpurun_out/r2z_ncu_hbm.log 2>&1; tail -2 gpurun_out/r2z_ncu_hbm.log
