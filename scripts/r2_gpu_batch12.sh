# round-2 GPU batch 12: tensor-core (3xTF32) GEMMs around the decoder: decoder / training tests, checkpoint render test, c2 + c4 bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "aggregate or train or fused or graphed or checkpoint or golden or tc" > gpurun_out/r2o_pytest.log 2>&1; tail -6 gpurun_out/r2o_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2o_c2.json 2> gpurun_out/r2o_c2.err; tail -2 gpurun_out/r2o_c2.err
python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2o_c4.json 2> gpurun_out/r2o_c4.err; tail -2 gpurun_out/r2o_c4.err
