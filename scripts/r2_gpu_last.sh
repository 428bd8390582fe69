# round-2 last batch: render lines with the visited-sample accounting, ncu --set full of composite_fwd / adam_multi
set -x
mkdir -p gpurun_out
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r2y_c1.json 2> gpurun_out/r2y_c1.err; tail -2 gpurun_out/r2y_c1.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2y_c3.json 2> gpurun_out/r2y_c3.err; tail -2 gpurun_out/r2y_c3.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_c5.json 2> gpurun_out/r2y_c5.err; tail -2 gpurun_out/r2y_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:composite_fwd_kernel|adam_multi_kernel" -c 4 -o gpurun_out/r2y_hbm_kernels -f python scripts/kernel_roofline.py > gpurun_out/r2y_ncu_hbm.log 2>&1; tail -2 gpurun_out/r2y_ncu_hbm.log
timeout 300 python -m pytest tests -q -m gpu -x -k "render or checkpoint or composite" 2>&1 | tail -2
