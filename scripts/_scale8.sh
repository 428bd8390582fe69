python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/s8_c2.json 2> gpurun_out/s8_c2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload c3 --no-cpu-baseline > gpurun_out/s8_c3.json 2> gpurun_out/s8_c3.err
tail -c 300 gpurun_out/s8_c2.err; tail -c 300 gpurun_out/s8_c3.err
python -c "
import json
for w in ('c2','c3'):
    try:
        d=json.loads(open('gpurun_out/s8_%s.json'%w).read().strip().splitlines()[-1]); print(w, d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))
    except Exception as e: print(w, 'ERR', e)
"
