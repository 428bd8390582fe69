# round-2 GPU batch 1: parity tests, first bench lines (c2, c4), k-NN timings + ncu captures of the shipped searches
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest.log 2>&1; tail -5 gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2a_c2.json 2> gpurun_out/r2a_c2.err; tail -c 600 gpurun_out/r2a_c2.json
python bench.py --workload c4 --no-cpu-baseline --stages > gpurun_out/r2a_c4.json 2> gpurun_out/r2a_c4.err; tail -c 600 gpurun_out/r2a_c4.json; tail -30 gpurun_out/r2a_c4.err
for w in c1 c3 c5; do python scripts/knn_profile.py $w > gpurun_out/r2_knn_times_$w.txt 2>&1; done
for w in c1 c3 c5; do
  python scripts/knn_profile.py $w short > gpurun_out/r2_knn_plain_$w.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:knn -c 2 -f -o gpurun_out/r2_knn_$w python scripts/knn_profile.py $w short > gpurun_out/r2_knn_ncu_$w.log 2>&1
done
ls -la gpurun_out | tail -20
