# round-2 GPU batch 7: sorted k-NN (exact float-box staging): bit-exact tests, then timings on c1 / c3 / c5 for a few first radii
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -m gpu -k "knn or sample" > gpurun_out/r2g_knn_tests.log 2>&1; tail -5 gpurun_out/r2g_knn_tests.log
rm -f gpurun_out/r2g_knn_sweep.txt
for cfg in "6 0.6" "6 0.4" "6 0.8" "6 1.0" "6 0.25"; do
  set -- $cfg
  for w in c1 c3 c5; do
    APN_KS_SUBBITS=$1 APN_KS_RHO0=$2 APN_KNN_FORCE=sorted timeout 300 python scripts/knn_profile.py $w time 2>&1 | tail -1 | sed "s/^/sub=$1 rho0=$2 /" | tee -a gpurun_out/r2g_knn_sweep.txt
  done
done
