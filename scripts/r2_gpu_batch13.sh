# round-2 GPU batch 13: sorted vs legacy k-NN inside the static (graph-captured) training step, c2 and c4
set -x
mkdir -p gpurun_out
for m in sorted legacy; do
  for w in c2 c4; do
    APN_KNN_STATIC=$m python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2p_${w}_$m.json 2> gpurun_out/r2p_${w}_$m.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_${w}_$m.json").read().strip().splitlines()[-1])
print("$w $m", round(d["ms_per_step"],3), "ms; sample+knn", d["stages_ms_per_step"].get("sample_ray+knn"))
PY
  done
done
