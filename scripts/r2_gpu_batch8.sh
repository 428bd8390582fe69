# round-2 GPU batch 8: launch lists of one sample_and_knn call (c1 / c3 / c5) and ncu --set full of the sorted search at c3 and c5
set -x
mkdir -p gpurun_out
for w in c1 c3 c5; do
  APN_KNN_FORCE=sorted timeout 300 python scripts/knn_profile.py $w short > gpurun_out/r2h_knn_plain_$w.log 2>&1 && \
  APN_KNN_FORCE=sorted timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_knn_launches_$w.csv python scripts/knn_profile.py $w short > /dev/null 2>&1
done
APN_KNN_FORCE=sorted timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_sorted_kernel -c 1 -o gpurun_out/r2h_knn_sorted_c3 -f python scripts/knn_profile.py c3 short > gpurun_out/r2h_knn_ncu_c3.log 2>&1
APN_KNN_FORCE=sorted timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_sorted_kernel -c 1 -o gpurun_out/r2h_knn_sorted_c5 -f python scripts/knn_profile.py c5 short > gpurun_out/r2h_knn_ncu_c5.log 2>&1
