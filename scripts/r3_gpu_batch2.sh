# backward fork after tc_dgrad (phase 3 | 4): tests + A/B
set -x
timeout 300 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "graphed or full or fused or derived or phases" > gpurun_out/r3c_pytest.log 2>&1; tail -3 gpurun_out/r3c_pytest.log
for i in 1 2 3; do
python bench.py --no-cpu-baseline --steps 40 > gpurun_out/r3c_c2_br_$i.json 2> gpurun_out/r3c_c2_br_$i.err
done
python bench.py --no-cpu-baseline --workload c4 --steps 40 > gpurun_out/r3c_c4_br.json 2> gpurun_out/r3c_c4_br.err
python bench.py --no-cpu-baseline --full-loss --steps 40 > gpurun_out/r3c_c2full_br.json 2> gpurun_out/r3c_c2full_br.err
