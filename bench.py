#!/usr/bin/env python
"""bench.py — throughput of the PCD hot path (BASELINE.json metric: rays/s, train fwd+bwd+Adam / render fwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4|c1|c3|c5] [--impl reference]

Default workload = BASELINE.json configs[1] ("c2"): jumpingjacks-shaped stage-2 training step, 8192-ray batch,
forward + render-loss backward + (masked) Adam over the reference's parameter groups.  Render workloads
(c1/c3/c5) time one whole frame per step.  Data: seeded synthetic scene of the reference's shapes
(articulated_point_nerf_b200/scene.py), random-init weights of the reference's architecture.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step through the public
Python surface with HOST (pinned) inputs, H2D copy + loss/frame D2H inside the timed region.
L2 is flushed (256 MiB write) between timed steps; every step is bracketed by CUDA events on the launching
stream; max over ranks.  `--impl reference` times the CPU port of the reference path (oracle/) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

TRAIN_WORKLOADS = {"c2": "c2", "c4": "c4", "tiny": "tiny", "small": "small"}
RENDER_WORKLOADS = {"c1": "c1", "c3": "c3", "c5": "c5", "tiny_render": "tiny", "small_render": "small"}
N_RAND = 8192                      # configs/nerf/default.py:114
METRIC = {"train": "rays/sec (PCD train step: fwd+bwd+Adam)", "render": "rays/sec (PCD render fwd)"}


# ----------------------------------------------------------------------------------------------
# workload construction (CPU, seeded): identical for both arms
# ----------------------------------------------------------------------------------------------
def make_batches(scene, mode, n_steps, rank, n_rand=N_RAND, repose=False):
    """Per step: (t, rays_o, rays_d, viewdirs, target) on the host.  Train: n_rand random pixels of one view
    (run.py:587-601: a D-NeRF time step is one image); render: every pixel of one view."""
    out = []
    n_views = len(scene.HW)
    cache = {}
    for i in range(n_steps):
        # train: every rank draws ITS pixels from the SAME view / time step (one iteration of run.py:587-601 is one time step: the
        # data-parallel batch is n_rand x ranks rays of it, sharded); render: a different view per rank (view-per-GPU)
        v = (i % n_views) if mode == "train" else (i + 3 * rank) % n_views
        if v not in cache:
            cache[v] = [x.reshape(-1, 3).contiguous() for x in scene.rays(v)]
        ro, rd, vd = cache[v]
        t = torch.tensor([v / max(n_views - 1, 1)], dtype=torch.float32)
        if repose:
            # run.py:1364-1377: random bone rotations, root fixed, scaled along the clip
            g = torch.Generator().manual_seed(77)
            rp = torch.randn(len(scene.joints), 4, generator=g) * 0.2
            rp[0] = 0
            t = (rp * ((i % 30) / 29.0)).contiguous()
        if mode == "train":
            g = torch.Generator().manual_seed(1000 * rank + i)
            sel = torch.randint(0, len(ro), (n_rand,), generator=g)
            tgt = torch.rand(n_rand, 3, generator=g)
            out.append((t, ro[sel].contiguous(), rd[sel].contiguous(), vd[sel].contiguous(), tgt))
        else:
            out.append((t, ro, rd, vd, None))
    return out


def pack_host(batch, pin):
    """One contiguous host buffer per step: [rays_o | rays_d | viewdirs | target] (R, 9 or 12) + t."""
    t, ro, rd, vd, tgt = batch
    parts = [ro, rd, vd] + ([tgt] if tgt is not None else [])
    buf = torch.cat(parts, dim=1).contiguous()
    if pin:
        buf = buf.pin_memory()
        t = t.pin_memory()
    return t, buf


def unpack_dev(t_dev, buf_dev):
    ro, rd, vd = buf_dev[:, 0:3].contiguous(), buf_dev[:, 3:6].contiguous(), buf_dev[:, 6:9].contiguous()
    tgt = buf_dev[:, 9:12].contiguous() if buf_dev.shape[1] >= 12 else None
    return t_dev, ro, rd, vd, tgt


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, n_gpus):
        """ONE sampler process for all GPUs of the job (rank 0 starts it): eight concurrent `nvidia-smi -lms` loops, one per
        rank, contend for the driver and slowed every rank's host-bound step by ~45 % on the 8-GPU run."""
        self.n, self.proc, self.path = n_gpus, None, f"/tmp/apn_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", ",".join(str(i) for i in range(self.n))],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                for n, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
            os.remove(self.path)
        except Exception:
            pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# algorithmic work of the decoder, from measured counts
# ----------------------------------------------------------------------------------------------
def decoder_flops(M, d_in, train):
    fwd = M * 8 * 2 * (d_in * 128 + 3 * 128 * 128) + M * 2 * (128 * 128 + 155 * 64 + 64 * 3 + 128)
    return fwd * (3 if train else 1)        # backward = dgrad + wgrad = 2x forward


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (oracle/), all host threads
# ----------------------------------------------------------------------------------------------
class OracleArm:
    def __init__(self, scene, model_state, mode):
        from oracle.path_oracle import OraclePath
        from oracle import dvgo_ops
        self.dvgo = dvgo_ops
        cfg = scene.cfg
        self.scene, self.cfg, self.mode = scene, cfg, mode
        state = {k: v.detach().cpu().clone() for k, v in model_state["state"].items()}
        self.orc = OraclePath(state, scene.canonical_pcd, scene.bones, stepsize=cfg.stepsize, voxel_size=scene.voxel_size,
                              fast_color_thres=cfg.fast_color_thres, act_shift=model_state["act_shift"],
                              voxel_size_ratio=model_state["voxel_size_ratio"], pose_embedding_dim=cfg.pose_embedding_dim,
                              mean_min_distance=model_state["mean_min_distance"])
        from articulated_point_nerf_b200.train import STAGE2_LRATES
        self.groups = []
        for name in ("rgbnet", "densitynet", "canonical_feat", "gammas", "weights", "theta_weight", "forward_warp", "joints",
                     "feat_net"):
            ks = [k for k in state if (k == name or k.startswith(name + ".")) and state[k].is_floating_point()]
            self.groups.append((STAGE2_LRATES[name], ks))
        self.trainable = [k for _, ks in self.groups for k in ks]
        self.adam = {k: (torch.zeros_like(state[k]), torch.zeros_like(state[k])) for k in self.trainable}
        self.step_no = 0

    def step(self, batch, n_rays=None):
        t, ro, rd, vd, tgt = batch
        if n_rays is not None and n_rays < len(ro):
            ro, rd, vd = ro[:n_rays], rd[:n_rays], vd[:n_rays]
            tgt = None if tgt is None else tgt[:n_rays]
        cfg, orc = self.cfg, self.orc
        kw = dict(rays_o=ro, rays_d=rd, viewdirs=vd, near=cfg.near, far=cfg.far, stepsize=cfg.stepsize, bg=cfg.bg)
        if self.mode == "render":
            with torch.no_grad():
                # the reference renders a frame in 8192-ray chunks, re-warping per chunk (run.py:136-166)
                outs = []
                for s in range(0, len(ro), N_RAND):
                    kwc = dict(kw, rays_o=ro[s:s + N_RAND], rays_d=rd[s:s + N_RAND], viewdirs=vd[s:s + N_RAND])
                    outs.append((orc.forward(None, t, **kwc) if t.dim() == 2 else orc.forward(t, **kwc))["rgb_marched"])
                return torch.cat(outs)
        for k in self.trainable:
            orc.s[k].requires_grad_(True)
            orc.s[k].grad = None
        out = orc.forward(t, **kw)
        loss = 200.0 * torch.nn.functional.mse_loss(out["rgb_marched"], tgt)
        if loss.requires_grad:
            loss.backward()
        self.step_no += 1
        with torch.no_grad():
            for lr, ks in self.groups:
                for k in ks:
                    p = orc.s[k]
                    if p.grad is None:
                        continue
                    m, v = self.adam[k]
                    self.dvgo.adam_upd(p, p.grad, m, v, self.step_no, 0.9, 0.99, lr, 1e-8)
        return loss.detach()


def time_oracle(arm, batches, steps, warmup, budget_s):
    """Bounded sample: rays per step are cut so that warmup+steps fit the budget (calibrated on one step)."""
    R = len(batches[0][1])
    t0 = time.perf_counter()
    arm.step(batches[0], min(R, 1024))
    probe = time.perf_counter() - t0
    # cost model: fixed (warp, tables) + per-ray; the 1024-ray probe over-estimates the per-ray part
    n = int(min(R, max(512, 1024 * (budget_s / max(steps + warmup, 1)) / max(probe, 1e-3))))
    for i in range(warmup):
        arm.step(batches[i % len(batches)], n)
    t0 = time.perf_counter()
    for i in range(steps):
        arm.step(batches[(warmup + i) % len(batches)], n)
    dt = time.perf_counter() - t0
    return n, dt / max(steps, 1)


# ----------------------------------------------------------------------------------------------
def model_state_for_oracle(model):
    return {"state": {k: v for k, v in model.state_dict().items() if not k.startswith("tineuvox.")},
            "act_shift": float(model.tineuvox.act_shift), "voxel_size_ratio": float(model.tineuvox.voxel_size_ratio),
            "mean_min_distance": None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--decoder", default="auto", choices=["auto", "fp32", "tc", "tc_fast"], help="inference decoder")
    ap.add_argument("--decoder-train", default="tc", choices=["tc", "fp32"], help="decoder of the training step")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds for the whole --impl reference run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stages", action="store_true", help="print the per-stage CUDA-event breakdown to stderr")
    ap.add_argument("--full-loss", action="store_true",
                    help="training workloads: the COMPLETE stage-2 iteration of run.py:615-694 — render loss + ARAP + weight TV + "
                         "sparsity + transformation regulariser + joint chamfer (configs/nerf/default.py:96-103) + the 2-D chamfer "
                         "term against synthetic mask pixels (5 views x 3000 pixels, 3000 random projected points)")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one all-reduce of the whole bucket after the backward (A/B)")
    ap.add_argument("--no-branches", action="store_true",
                    help="graphed step on ONE stream: no side-stream branch for the decoder state / the early Adam part (A/B)")
    ap.add_argument("--train-path", default="graph", choices=["graph", "static", "dynamic"],
                    help="training step: CUDA graphs over the sync-free step (default), the same step launched eagerly, "
                         "or the dynamic step with its two host read-backs")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: at least 3 warm-up steps"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    mode = "train" if args.workload in TRAIN_WORKLOADS else "render"
    cfg_name = TRAIN_WORKLOADS.get(args.workload) or RENDER_WORKLOADS[args.workload]

    from articulated_point_nerf_b200.scene import build_model, make_scene
    scene = make_scene(cfg_name)
    repose = mode == "render" and cfg_name in ("c3", "c5")          # --repose_pcd workloads (run.py:1355-1396)
    n_steps = args.steps + args.warmup
    base_cfg = {"workload": f"{args.workload}: {mode}, N={len(scene.canonical_pcd)} points, J={len(scene.joints)}, "
                            f"{'%d-ray batch of one %dx%d view' % (N_RAND, scene.cfg.H, scene.cfg.W) if mode == 'train' else 'one %dx%d frame' % (scene.cfg.H, scene.cfg.W)} per step per GPU",
                "l2": "flushed (256 MiB write) between timed steps", "parallelism": f"rays sharded, dp{args.gpus}",
                "train_path": args.train_path if mode == "train" else None}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        model = build_model(scene, seed=0)
        arm = OracleArm(scene, model_state_for_oracle(model), mode)
        batches = make_batches(scene, mode, min(n_steps, 8), 0, repose=repose)
        n, sec = time_oracle(arm, batches, args.steps, args.warmup, args.ref_budget)
        val = n / sec
        sample = f"{n} of {len(batches[0][1])} rays per step, {args.steps} steps, oracle port (torch-CPU fp32, brute-force k-NN)"
        print(json.dumps({
            "impl": "reference", "metric": METRIC[mode], "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": base_cfg,
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ B200 arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    from articulated_point_nerf_b200 import _lib
    from articulated_point_nerf_b200.train import GradBucket, GraphedTrainStep, create_optimizer, make_bucket, train_step

    model = build_model(scene, seed=0)
    oracle_state = model_state_for_oracle(model) if (rank == 0 and not args.no_cpu_baseline and world == 1) else None
    if oracle_state is not None:
        oracle_state["state"] = {k: v.detach().clone() for k, v in oracle_state["state"].items()}
    model = model.to(dev)
    if os.environ.get("APN_NO_POSE_GRAPH"):
        model.graph_pose = False
    if args.decoder != "auto":
        model.decoder = args.decoder
    model.decoder_train = args.decoder_train
    host = [pack_host(b, pin=True) for b in make_batches(scene, mode, n_steps, rank, repose=repose)]
    n_views = len(scene.HW)
    views = [(i + 3 * rank) % n_views for i in range(n_steps)]        # render: the view of step i (as make_batches picks it)
    rk = scene.render_kwargs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    opt = bucket = None
    if mode == "train":
        opt = create_optimizer(model)
        # default: the pipelined layout (decoder slice reduced beside the next step's sampling stage); --no-overlap: one all-reduce
        bucket = GradBucket(opt) if args.no_overlap else make_bucket(model, opt)
    decay = 0.1 ** (1.0 / (160 * 1000))
    counts = []
    gs = gs_stages = None
    reg = extra = None
    if mode == "train" and args.full_loss:
        from articulated_point_nerf_b200.train import Chamfer2D, Regularisers
        reg = Regularisers()
        gen = torch.Generator().manual_seed(123)
        Bv = min(5, len(scene.poses))                     # run.py:663: at most 5 cameras of the time step
        mask_pcd = torch.stack([torch.randint(0, scene.cfg.H, (Bv, 3000), generator=gen),
                                torch.randint(0, scene.cfg.W, (Bv, 3000), generator=gen)], dim=-1).float()
        extra = Chamfer2D(model, scene.poses[:Bv].float().to(dev), scene.Ks[:Bv].float().to(dev), mask_pcd.to(dev), weight=5e-3,
                          n_points=3000, image_height=None if scene.cfg.inverse_y else scene.cfg.H)
        base_cfg["loss"] = "render + arap + weight_tv + sparsity + transformation_reg + joint_chamfer + chamfer2D (run.py:615-694)"
    elif mode == "train":
        base_cfg["loss"] = "render loss only (run.py:615-631)"
    if mode == "train" and args.train_path != "dynamic":
        t0, b0 = host[0][0].to(dev), host[0][1].to(dev)
        cal = (t0, b0[:, 0:3].contiguous(), b0[:, 3:6].contiguous())
        gs = GraphedTrainStep(model, opt, bucket, len(b0), rk, calibrate=cal, use_graph=args.train_path == "graph",
                              packed_inputs=True, regularisers=reg, extra_loss=extra)
        gs.branches = not args.no_branches
        base_cfg["graph_branches"] = gs.branches
        if world > 1:
            base_cfg["grad_exchange"] = ("pipelined across steps (decoder slice beside the next step's sampling stage)" if gs.pipelined
                                         else "three parts beside the backward" if bucket.split > 0 else "one all-reduce after the backward")
        seen = set()
        for t_h, b_h in host:                       # one sampling pass per distinct view: the workspace fits the densest one
            key = float(t_h.reshape(-1)[0])
            if key not in seen:
                seen.add(key)
                b_d = b_h.to(dev)
                gs.reserve((t_h.to(dev), b_d[:, 0:3].contiguous(), b_d[:, 3:6].contiguous()))

    def run_step(t_dev, buf_dev, stepper=None, view=None):
        if mode == "train" and (stepper or gs) is not None:
            return (stepper or gs).step_packed(t_dev, buf_dev, decay)       # counts arrive later (gs.history)
        if view is not None:
            # render, end to end: the frame's rays come from the CAMERA (K, c2w: ~100 bytes, kernel arguments) in one launch
            # on the device (apn_rays_of_a_view) — what render.render_viewpoints does — instead of 36 bytes per pixel of H2D
            from articulated_point_nerf_b200 import ops as _ops
            t, tgt = t_dev, None
            ro, rd, vd = _ops.rays_of_a_view(scene.cfg.H, scene.cfg.W, scene.Ks[view], scene.poses[view], dev,
                                             inverse_y=scene.cfg.inverse_y)
        else:
            t, ro, rd, vd, tgt = unpack_dev(t_dev, buf_dev)
        kw = dict(rk, rays_o=ro, rays_d=rd, viewdirs=vd)
        if mode == "train":
            out = train_step(model, opt, bucket, t, kw, tgt, decay_factor=decay, regularisers=reg, extra_loss=extra)
        else:
            with torch.no_grad():
                if t.dim() == 2:      # repose: rot_params instead of a time (run.py:287)
                    out = model(None, render_depth=True, render_kwargs=kw, rot_params=t)["rgb_marched"]
                else:
                    out = model(t, render_depth=True, render_kwargs=kw)["rgb_marched"]
        counts.append(dict(model.last_counts))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(e2e: bool, stages: bool = False):
        """value loop (e2e=False): resident inputs, NO stage timers.  e2e loop: pinned host inputs, H2D + D2H inside.
        stages=True (a third, separate pass): resident inputs with the per-stage CUDA-event brackets on — it feeds
        `stages_ms_per_step` and the roofline figures, never `value`."""
        dev_in = None if e2e else [(t.to(dev), b.to(dev)) for t, b in host]
        stepper = None
        if stages and gs is not None:          # stage brackets are CUDA events: they cannot live inside a captured graph
            nonlocal gs_stages
            if gs_stages is None:
                gs_stages = GraphedTrainStep(model, opt, bucket, gs.R, rk, cand_cap=gs.cand_cap, m_cap=gs.m_cap, use_graph=False,
                                             packed_inputs=True, regularisers=reg, extra_loss=extra)
                gs_stages.branches = False       # stage brackets time a serial chain
            stepper = gs_stages
        on_host = e2e and (gs is not None)     # the graphed step takes the pinned host buffers directly (one H2D copy each)
        for i in range(args.warmup):
            cam = views[i] if (e2e and mode == "render") else None
            if on_host:
                t_d, b_d = host[i]
            elif cam is not None:
                t_d, b_d = host[i][0].to(dev, non_blocking=True), None
            else:
                t_d, b_d = (host[i][0].to(dev, non_blocking=True), host[i][1].to(dev, non_blocking=True)) if e2e else dev_in[i]
            o = run_step(t_d, b_d, stepper, view=cam)
            if e2e:
                (o.item() if mode == "train" else o.cpu())
        for st_ in (gs, gs_stages):
            if st_ is not None:
                st_.flush()
                st_.history.clear()
        counts.clear()
        if not e2e:
            # workspace headroom: sample counts differ from view to view, so later steps can need somewhat larger
            # buffers than any warm-up step; keep a cached block for the allocator to carve them from instead of
            # reaching cudaMalloc (which synchronises) inside the timed region
            spare = torch.empty(max(int(0.5 * torch.cuda.memory_reserved(dev)), 64 << 20), dtype=torch.uint8, device=dev)
            del spare
        barrier()
        _lib.STAGES.reset(stages)
        evs = []
        n0 = _lib.launch_count()
        # graphed training step: each step's loss travels to pinned host memory with the step; the host reads it one step
        # late (after the next step is enqueued), so every step's result IS read inside the timed region without draining
        # the GPU between steps
        lagged = e2e and mode == "train" and (stepper or gs) is not None
        pending_read = None
        t_wall = time.perf_counter()
        for i in range(args.warmup, n_steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            cam = views[i] if (e2e and mode == "render") else None
            if on_host:
                t_d, b_d = host[i]
            elif cam is not None:
                t_d, b_d = host[i][0].to(dev, non_blocking=True), None     # the pose (time or rot_params); the camera goes by value
            elif e2e:
                t_d, b_d = host[i][0].to(dev, non_blocking=True), host[i][1].to(dev, non_blocking=True)
            else:
                t_d, b_d = dev_in[i]
            o = run_step(t_d, b_d, stepper, view=cam)
            if lagged:
                if pending_read is not None:
                    pending_read()
                pending_read = (stepper or gs).loss_reader()
            elif e2e:
                (o.item() if mode == "train" else o.cpu())
            b.record()
            evs.append((a, b))
        if pending_read is not None:
            pending_read()
        barrier()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
        active = stepper or gs
        if active is not None:
            active.flush()
            counts.extend(dict(R=active.R, candidates=c_, M=m_) for c_, m_ in active.history)
            active.history.clear()
        if os.environ.get("APN_ALLOC_STATS") and rank == 0:
            st = torch.cuda.memory_stats()
            print(f"  alloc stats ({'e2e' if e2e else 'value'}): device_alloc={st.get('num_device_alloc')} "
                  f"device_free={st.get('num_device_free')} retries={st.get('num_alloc_retries')} "
                  f"reserved={st.get('reserved_bytes.all.current', 0) / 1e6:.0f} MB", file=sys.stderr)
        launches = _lib.launch_count() - n0
        if gs is not None and gs.graphs is not None and not stages:
            launches += (n_steps - args.warmup) * gs.launches_per_step       # kernels replayed from the graphs
        per_step = torch.tensor([a.elapsed_time(b) for a, b in evs], device=dev, dtype=torch.float64)
        tt = per_step.sum().reshape(1)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(per_step, op=dist.ReduceOp.MAX)          # per step: the slowest rank
        q = torch.quantile(per_step, torch.tensor([0.1, 0.5, 0.9], device=dev, dtype=torch.float64)).tolist()
        return float(tt.item()), launches, {"p10": q[0], "median": q[1], "p90": q[2], "min": float(per_step.min()),
                                            "max": float(per_step.max()),
                                            "host_wall_incl_l2_flush": wall_ms / max(n_steps - args.warmup, 1)}

    clocks = ClockSampler(world) if rank == 0 else None
    if clocks is not None:
        clocks.start()
    total_ms, launches, dist_ms = timed_loop(e2e=False)
    clk = clocks.stop() if clocks is not None else None
    e2e_ms, _, e2e_dist = timed_loop(e2e=True)
    _, _, _ = timed_loop(e2e=False, stages=True)            # separate pass: stage brackets cost ~22 event pairs per step
    stage_tot = _lib.STAGES.totals()
    _lib.STAGES.reset(False)
    timed_counts = list(counts)

    rays_per_step = len(host[0][1])
    value = rays_per_step * world * args.steps / (total_ms * 1e-3)
    e2e_val = rays_per_step * world * args.steps / (e2e_ms * 1e-3)
    # train: the packed (R, 12) batch + t.  render: the pose (t or rot_params) + the camera (K 3x3, c2w 3x4 as kernel arguments)
    h2d = (host[0][1].numel() * 4 + 4) if mode == "train" else (host[0][0].numel() * 4 + (9 + 12) * 4)
    d2h = 4 if mode == "train" else rays_per_step * 3 * 4

    # roofline of the dominant kernel group: the decoder (feat_net + heads; fwd [+ bwd]) — the only dense contraction
    M_sum = sum(c.get("M", 0) for c in timed_counts[:args.steps])
    d_in = 191 + scene.cfg.pose_embedding_dim
    dec_ms = sum(ms for name, (n, ms) in stage_tot.items() if name.startswith("feat_net"))
    dec_ms_timed = dec_ms          # stage events cover exactly the timed steps
    flops = decoder_flops(M_sum, d_in, mode == "train")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    ach_tf = flops / max(dec_ms_timed * 1e-3, 1e-9) / 1e12
    roofline = {"kernel": "decoder (feat_net + heads%s), %s path" % (" fwd+bwd" if mode == "train" else "",
                                                                      model.decoder_train if mode == "train" else model.decoder),
                "bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "traffic": None, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else
                "fallback 1.4 PFLOP/s sustained (of fallback)",
                "decoder_ms_per_step": dec_ms_timed / args.steps, "kept_samples_per_step": M_sum / max(args.steps, 1)}
    # measured DRAM traffic of the decoder kernels per step (one ncu --set full capture per round, profiles/)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tr:
            roofline["traffic"] = tr["dram_bytes_per_step"]
            roofline["traffic_source"] = tr["source"]
    except Exception:
        pass
    # the bandwidth-bound stages of the same timed steps against the HBM roofline: algorithmic bytes
    # from the measured counts / CUDA-event stage time.  At c1/c2 sizes these stages are latency-bound (a few MB per
    # launch); profiles/r01_kernel_roofline.json holds the same kernels at the 1M-point scale.
    N_pts, J_b, R_st = len(scene.canonical_pcd), len(scene.joints), rays_per_step
    M_avg = M_sum / max(args.steps, 1)
    T_avg = sum(c.get("candidates", 0) for c in timed_counts[:args.steps]) / max(args.steps, 1)
    # LBS forward: SURVEY §8(d) bytes N(4J+12+12+36)+64J; the merged weights (another 4J per point) are only written
    # when a caller needs them (training: the regularisers read _last_weights; no-grad renders skip the store)
    alg = {"forward_warp": N_pts * (4 * J_b + 12 + 12 + 36) + 64 * J_b,
           "forward_warp_bwd": N_pts * (8 * J_b + 12 + 12 + 36 + 36),
           "grid_build": N_pts * 32,
           "sample_ray+knn": R_st * 24 + T_avg + M_avg * (12 + 8 + 64) + 12 * N_pts,
           "Alphas2Weights": 2 * (M_avg * 24 + R_st * 24),          # main + direct branch
           "Alphas2Weights_bwd": M_avg * 40 + R_st * 24,
           "adam": 28 * sum(p.numel() for p in model.parameters() if p.requires_grad)}
    visited_note = None
    if mode == "render":
        # the compositing forward stops LOADING a ray's samples at the early stop (T < 1e-3): count the samples the walk
        # actually visits (one extra, untimed frame with ops.RECORD_VISITED) and charge only those
        from articulated_point_nerf_b200 import ops as _ops
        _ops.RECORD_VISITED = []
        t_d, b_d = host[n_steps - 1][0].to(dev), host[n_steps - 1][1].to(dev)
        run_step(t_d, b_d)
        torch.cuda.synchronize()
        visited = sum(int(v.sum().item()) for v in _ops.RECORD_VISITED)            # main + direct branch
        _ops.RECORD_VISITED = None
        M_last = counts[-1].get("M", 0) if counts else 0
        alg["Alphas2Weights"] = visited * 24 + 2 * R_st * 24
        visited_note = (f"bytes of the samples the walk visits before the early stop: {visited} of {2 * M_last} (main + direct branch) "
                        f"on the last frame")
    hbm_peak = peaks.get("hbm_gbs", 6500.0)
    hbm_stages = {}
    for name, (n, ms) in stage_tot.items():
        if name in alg and ms > 0:
            gbs = alg[name] * args.steps / (ms * 1e-3) / 1e9
            hbm_stages[name] = {"algorithmic_bytes_per_step": int(alg[name]), "achieved_gbs": round(gbs, 1),
                                "frac": round(gbs / hbm_peak, 4)}
            if name == "Alphas2Weights" and visited_note:
                hbm_stages[name]["note"] = visited_note
    if args.stages and rank == 0:
        for name, (n, ms) in sorted(stage_tot.items(), key=lambda kv: -kv[1][1]):
            print(f"  stage {name:22s} calls {n:4d}  total {ms:9.3f} ms  avg {ms / n:8.4f} ms", file=sys.stderr)

    cpu_baseline = None
    if oracle_state is not None:
        torch.set_num_threads(os.cpu_count() or 1)
        arm = OracleArm(scene, oracle_state, mode)
        batches = make_batches(scene, mode, 3, 0, repose=repose)
        n, sec = time_oracle(arm, batches, 2, 0, args.cpu_budget)
        cpu_baseline = {"value": n / sec, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} of {len(batches[0][1])} rays per step, 2 steps after a 1024-ray probe, oracle port "
                                  f"(torch-CPU fp32, brute-force k-NN) of the same workload"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC[mode], "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": base_cfg, "clocks": clk,
            "ms_per_step_dist": dist_ms,
            "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "ms_per_step_dist": e2e_dist},
            "gpu_launches": launches, "roofline": roofline, "roofline_hbm_stages": hbm_stages, "cpu_baseline": cpu_baseline,
            "counts": {"kept_samples_per_step": M_sum / max(args.steps, 1),
                       "candidates_per_step": sum(c.get("candidates", 0) for c in timed_counts[:args.steps]) / max(args.steps, 1)},
            "stages_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in stage_tot.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
