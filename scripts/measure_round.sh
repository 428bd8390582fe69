# Measurement batch behind profiles/r01_*_final.* (run on the B200 box: gpurun -- bash scripts/measure_round.sh).
# Every command that runs under ncu is first run without it, as the profiling recipe asks.
set -x
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/f_c2.json 2> gpurun_out/f_c2.err
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/f_c1.json 2> gpurun_out/f_c1.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/f_c3.json 2> gpurun_out/f_c3.err
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_c2_short.json 2> gpurun_out/f_c2_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/f_launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_c2.log 2>&1
python scripts/kernel_roofline.py > gpurun_out/f_roof.json 2> gpurun_out/f_roof.err && \
ncu --set full --clock-control none --import-source on -k regex:lbs_bwd_mma -c 1 -o gpurun_out/f_lbs_bwd -f python scripts/kernel_roofline.py > gpurun_out/f_ncu_lbs.log 2>&1
