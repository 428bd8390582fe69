# multi-GPU bench lines: bash scripts/r2_gpu_scale.sh N tag workloads...   (run under gpurun --gpus N)
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
for w in "$@"; do
  extra=""
  case $w in c5) extra="--steps 5 --warmup 3";; esac
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --no-cpu-baseline $extra $EXTRA > gpurun_out/${TAG}_${w}_${N}gpu.json 2> gpurun_out/${TAG}_${w}_${N}gpu.err
  tail -2 gpurun_out/${TAG}_${w}_${N}gpu.err; head -c 600 gpurun_out/${TAG}_${w}_${N}gpu.json; echo
done
