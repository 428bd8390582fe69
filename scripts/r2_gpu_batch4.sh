set -x
mkdir -p gpurun_out
for w in c1 c3 c5; do
  APN_KNN_FORCE=sorted timeout 300 python scripts/knn_profile.py $w time > gpurun_out/r2d_knn_${w}_sorted.txt 2>&1; tail -1 gpurun_out/r2d_knn_${w}_sorted.txt
done
APN_KNN_FORCE=sorted timeout 300 python scripts/knn_profile.py c2 time > gpurun_out/r2d_knn_c2_sorted.txt 2>&1; tail -1 gpurun_out/r2d_knn_c2_sorted.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; tail -12 gpurun_out/r2d_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2d_c2.json 2> gpurun_out/r2d_c2.err; tail -5 gpurun_out/r2d_c2.err
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r2d_c1.json 2> gpurun_out/r2d_c1.err; tail -5 gpurun_out/r2d_c1.err
