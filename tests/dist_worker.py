"""Worker of tests/test_gpu_dist.py: one rank of a 2-process data-parallel training step on ONE GPU (gloo carries the CUDA
tensors through the host; no kernel of one rank ever waits for another rank's kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(rank, world, port, mode, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from conftest import load_golden, model_from_golden
    from articulated_point_nerf_b200.train import (FusedTrainStep, GraphedTrainStep, create_optimizer, make_bucket, shard_rays)
    g = load_golden("tiny")
    model, scene = model_from_golden(g, fused_pose=True)
    model.decoder_train = "tc"
    opt = create_optimizer(model)
    # graph1: ONE all-reduce of the whole bucket per step; pipe*: the exchange pipelined across steps (the default bucket);
    # static / graph: three parts beside the backward of the same step
    bucket = make_bucket(model, opt, overlap={"graph1": False, "pipe": "pipeline", "pipe_static": "pipeline"}.get(mode, True))
    R = len(g["rays_o"])
    a, b = shard_rays(R, rank, world)
    t = g["train"]["t"].cuda()
    ro, rd, vd, tgt = (x[a:b].contiguous().cuda() for x in (g["rays_o"], g["rays_d"], g["viewdirs"], g["train"]["target"]))
    rk = dict(scene.render_kwargs(), rays_o=ro, rays_d=rd, viewdirs=vd)
    if mode == "grads":
        # the gradient exchange alone: fused forward + backward on this rank's shard, then the bucket all-reduce
        with bucket.direct_accum():
            loss = FusedTrainStep(model, opt, bucket).run(t, rk, tgt)
        bucket.all_reduce_avg()
        torch.cuda.synchronize()
        out = {"flat": bucket.flat[:bucket.total].cpu(), "loss": float(loss), "M": model.last_counts["M"]}
    else:
        # the graph-captured step with its split all-reduce (early slice on the communication stream), two iterations
        before = {k: p.detach().clone() for k, p in model.named_parameters()}
        gs = GraphedTrainStep(model, opt, bucket, b - a, scene.render_kwargs(), calibrate=(t, ro, rd), use_graph=not mode.endswith("static"))
        assert gs.pipelined == mode.startswith("pipe")
        losses = [float(gs.step(t, ro, rd, vd, tgt)) for _ in range(3)]
        gs.flush()
        out = {"params": {k: p.detach().cpu() for k, p in model.named_parameters()}, "losses": losses, "M": gs.last_counts["M"],
               "split": bucket.split, "total": bucket.total,
               "moved": sorted(k for k, p in model.named_parameters() if not torch.equal(p.detach(), before[k]))}
    torch.save(out, os.path.join(out_dir, f"rank{rank}_{mode}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
