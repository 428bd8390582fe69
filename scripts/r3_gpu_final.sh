# final check of the round: the whole GPU suite, then the default bench line (c2, with the CPU baseline leg)
set -x
timeout 200 python -m pytest tests -q -m gpu > gpurun_out/r3z_pytest.log 2>&1; tail -4 gpurun_out/r3z_pytest.log
timeout 100 python bench.py --cpu-budget 12 > gpurun_out/r3z_c2.json 2> gpurun_out/r3z_c2.err; tail -c 300 gpurun_out/r3z_c2.json
