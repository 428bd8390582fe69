# round-2 GPU batch 5: whole suite, bench lines c2/c4/c1/c3/c5, c2 launch list, kernel roofline, ncu --set full of the sorted k-NN (c3)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; tail -12 gpurun_out/r2e_pytest.log
python bench.py > gpurun_out/r2e_c2.json 2> gpurun_out/r2e_c2.err; tail -3 gpurun_out/r2e_c2.err
python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2e_c4.json 2> gpurun_out/r2e_c4.err; tail -3 gpurun_out/r2e_c4.err
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r2e_c1.json 2> gpurun_out/r2e_c1.err; tail -3 gpurun_out/r2e_c1.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2e_c3.json 2> gpurun_out/r2e_c3.err; tail -3 gpurun_out/r2e_c3.err
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_c5.json 2> gpurun_out/r2e_c5.err; tail -3 gpurun_out/r2e_c5.err
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --train-path static > gpurun_out/r2e_c2_short.json 2> gpurun_out/r2e_c2_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2e_launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --train-path static > gpurun_out/r2e_ncu_c2.log 2>&1
python scripts/kernel_roofline.py > gpurun_out/r2e_roof.json 2> gpurun_out/r2e_roof.err; tail -3 gpurun_out/r2e_roof.err
APN_KNN_FORCE=sorted timeout 300 python scripts/knn_profile.py c3 short > gpurun_out/r2e_knn_plain.log 2>&1 && \
APN_KNN_FORCE=sorted timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_sorted_kernel -c 1 -o gpurun_out/r2e_knn_sorted_c3 -f python scripts/knn_profile.py c3 short > gpurun_out/r2e_knn_ncu.log 2>&1
ls -la gpurun_out | tail -5
