# round-2 GPU batch 3: k-NN searches (bit-exact tests + timings), then the whole suite and c2 / c1 / c3 benches
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -m gpu -k "knn or sample" > gpurun_out/r2c_knn_tests.log 2>&1; tail -15 gpurun_out/r2c_knn_tests.log
for w in c1 c3 c5; do
  for a in sorted legacy; do
    if [ $a = sorted ]; then export APN_KNN_FORCE=sorted; else unset APN_KNN_FORCE; export APN_KNN_LEGACY=1; fi
    timeout 300 python scripts/knn_profile.py $w time > gpurun_out/r2c_knn_${w}_$a.txt 2>&1; tail -3 gpurun_out/r2c_knn_${w}_$a.txt
    unset APN_KNN_FORCE APN_KNN_LEGACY
  done
done
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; tail -15 gpurun_out/r2c_pytest.log
APN_KNN_STATIC=legacy python bench.py --no-cpu-baseline > gpurun_out/r2c_c2_legacyknn.json 2> gpurun_out/r2c_c2_legacyknn.err
python bench.py --no-cpu-baseline > gpurun_out/r2c_c2.json 2> gpurun_out/r2c_c2.err; tail -5 gpurun_out/r2c_c2.err
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r2c_c1.json 2> gpurun_out/r2c_c1.err; tail -5 gpurun_out/r2c_c1.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2c_c3.json 2> gpurun_out/r2c_c3.err; tail -5 gpurun_out/r2c_c3.err
