# round-2 GPU batch 9b: sorted k-NN with coarse-level enumeration: timings (no counters)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2i_knn_sweep2.txt
for lv in "1e9,1e9" "1.5,1e9" "1.0,1e9" "0.5,1e9"; do
  for w in c1 c3 c5; do
    APN_KS_LEVELS=$lv APN_KNN_FORCE=sorted timeout 120 python scripts/knn_profile.py $w time 2>&1 | tail -1 | sed "s/^/levels=$lv /" | tee -a gpurun_out/r2i_knn_sweep2.txt
  done
done
for gr in "1,2" "2,0" "1.5,1"; do
  for w in c1 c3 c5; do
    APN_KS_GROWTH=$gr APN_KS_LEVELS=1.0,1e9 APN_KNN_FORCE=sorted timeout 120 python scripts/knn_profile.py $w time 2>&1 | tail -1 | sed "s/^/growth=$gr /" | tee -a gpurun_out/r2i_knn_sweep2.txt
  done
done
