# graph branches (decoder state beside the sampling chain, early Adam part beside the warp backward): tests + A/B
set -x
timeout 300 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "graphed or full or fused or derived" > gpurun_out/r3a_pytest.log 2>&1; tail -3 gpurun_out/r3a_pytest.log
timeout 200 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r3a_pytest_dist.log 2>&1; tail -2 gpurun_out/r3a_pytest_dist.log
for i in 1 2; do
python bench.py --no-cpu-baseline --steps 40 > gpurun_out/r3a_c2_br_$i.json 2> gpurun_out/r3a_c2_br_$i.err
python bench.py --no-cpu-baseline --steps 40 --no-branches > gpurun_out/r3a_c2_nobr_$i.json 2> gpurun_out/r3a_c2_nobr_$i.err
done
python bench.py --no-cpu-baseline --workload c4 --steps 40 > gpurun_out/r3a_c4_br.json 2> gpurun_out/r3a_c4_br.err
python bench.py --no-cpu-baseline --workload c4 --steps 40 --no-branches > gpurun_out/r3a_c4_nobr.json 2> gpurun_out/r3a_c4_nobr.err
python bench.py --no-cpu-baseline --full-loss --steps 40 > gpurun_out/r3a_c2full_br.json 2> gpurun_out/r3a_c2full_br.err
python bench.py --no-cpu-baseline --full-loss --steps 40 --no-branches > gpurun_out/r3a_c2full_nobr.json 2> gpurun_out/r3a_c2full_nobr.err
