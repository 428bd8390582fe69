# round-2 GPU batch 2: parity tests incl. reference kernels + static/graphed step; c2 bench in the three training paths; c4
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; tail -25 gpurun_out/r2b_pytest.log
for p in graph static dynamic; do
  python bench.py --train-path $p --no-cpu-baseline > gpurun_out/r2b_c2_$p.json 2> gpurun_out/r2b_c2_$p.err; tail -c 300 gpurun_out/r2b_c2_$p.err
done
python bench.py --workload c4 --no-cpu-baseline --stages > gpurun_out/r2b_c4.json 2> gpurun_out/r2b_c4.err; tail -20 gpurun_out/r2b_c4.err
