"""CPU: size-independent properties of the oracle's restated ops (oracle/dvgo_ops.py, oracle/path_oracle.py) — each op against
the most literal scalar form of the reference kernel it restates, against autograd of its closed form, and against the
invariants of the domain (telescoping transmittance, sortedness, tie-break order).  Complements test_oracle_vs_reference.py,
which pins the same functions against the reference's own run."""
import math

import numpy as np
import torch

from oracle import dvgo_ops
from oracle.path_oracle import knn_bruteforce


def _ragged_rays(gen, n_rays, max_len, p_empty=0.2):
    lens = torch.randint(0, max_len + 1, (n_rays,), generator=gen)
    lens[torch.rand(n_rays, generator=gen) < p_empty] = 0
    ray_id = torch.repeat_interleave(torch.arange(n_rays), lens)
    return lens, ray_id


def _alpha2weight_scalar(alpha, ray_id, n_rays):
    """lib/cuda/render_utils_kernel.cu:431-459 read literally: one ray after the other, float T_cum updated through a double
    product, early stop below 1e-3."""
    w = np.zeros(len(alpha), np.float32)
    T = np.ones(len(alpha), np.float32)
    last = np.ones(n_rays, np.float32)
    a = alpha.numpy()
    rid = ray_id.numpy()
    i = 0
    while i < len(a):
        r = rid[i]
        t = np.float32(1.0)
        j = i
        stopped = False
        while j < len(a) and rid[j] == r:
            if not stopped:
                T[j] = t
                w[j] = t * a[j]
                t = np.float32(np.float64(t) * (1.0 - np.float64(a[j])))
                if np.float64(t) < 1e-3:
                    stopped = True
            j += 1
        last[r] = t
        i = j
    return w, T, last


def test_alpha2weight_equals_the_scalar_kernel_and_telescopes():
    gen = torch.Generator().manual_seed(0)
    for n_rays, max_len, hi in [(1, 5, 0.5), (37, 40, 0.3), (64, 200, 0.9), (5, 0, 0.5)]:
        lens, ray_id = _ragged_rays(gen, n_rays, max_len)
        alpha = torch.rand(len(ray_id), generator=gen) * hi
        w, T, last, i_start, i_end = dvgo_ops.alpha2weight(alpha, ray_id, n_rays)
        ws, Ts, ls = _alpha2weight_scalar(alpha, ray_id, n_rays)
        assert np.array_equal(w.numpy(), ws)
        assert np.array_equal(last.numpy(), ls)
        visited = w.numpy() != 0
        assert np.array_equal(T.numpy()[visited], Ts[visited])
        # transmittance telescopes: sum of a ray's weights + what is left behind it == 1 (also for rays that stop early)
        tot = torch.zeros(n_rays).index_add_(0, ray_id, w) + last
        assert float((tot - 1).abs().max()) < 2e-6 if n_rays else True
        assert bool((last[lens == 0] == 1).all())
        # early termination: no weight behind the first sample whose running transmittance fell below 1e-3
        for r in range(n_rays):
            seg = slice(int(i_start[r]), int(i_start[r]) + int(lens[r]))
            t = np.cumprod(1.0 - alpha[seg].double().numpy())
            stop = np.nonzero(t < 1e-3 * (1 - 1e-6))[0]
            if len(stop):
                assert float(w[seg][stop[0] + 1:].abs().sum()) == 0.0


def test_alpha2weight_backward_equals_autograd_of_the_closed_form():
    gen = torch.Generator().manual_seed(1)
    n_rays = 23
    lens, ray_id = _ragged_rays(gen, n_rays, 30)
    alpha = (torch.rand(len(ray_id), generator=gen) * 0.15).double()        # no ray reaches the 1e-3 stop
    gw = torch.randn(len(ray_id), generator=gen).double()
    gl = torch.randn(n_rays, generator=gen).double()
    a = alpha.clone().requires_grad_(True)
    loss = 0
    o = 0
    for r in range(n_rays):
        n = int(lens[r])
        seg = a[o:o + n]
        T = torch.cat([torch.ones(1, dtype=torch.float64), torch.cumprod(1 - seg, 0)])
        loss = loss + (T[:-1] * seg * gw[o:o + n]).sum() + T[-1] * gl[r]
        o += n
    loss.backward()
    w, T, last, i_start, i_end = dvgo_ops.alpha2weight(alpha.float(), ray_id, n_rays)
    g = dvgo_ops.alpha2weight_backward(alpha.float(), w, T, last, i_start, i_end, n_rays, gw.float(), gl.float())
    assert float((g.double() - a.grad).abs().max()) < 1e-5 * float(a.grad.abs().max())


def test_raw2alpha_and_its_backward_equal_the_closed_form():
    gen = torch.Generator().manual_seed(2)
    d = (torch.randn(4096, generator=gen) * 4).double().requires_grad_(True)
    shift, interval = -6.9, 0.5
    alpha_ref = 1 - torch.pow(1 + torch.exp(d + shift), -interval)
    gb = torch.randn(4096, generator=gen).double()
    (alpha_ref * gb).sum().backward()
    e, alpha = dvgo_ops.raw2alpha(d.detach().float(), shift, interval)
    assert float((alpha.double() - alpha_ref.detach()).abs().max()) < 1e-6
    g = dvgo_ops.raw2alpha_backward(e, gb.float(), interval)
    assert float((g.double() - d.grad).abs().max()) < 1e-5 * float(d.grad.abs().max())


def test_sample_pts_on_rays_invariants():
    gen = torch.Generator().manual_seed(3)
    n = 300
    o = torch.randn(n, 3, generator=gen) * 0.2 + torch.tensor([0.0, 0.0, -3.0])
    d = torch.randn(n, 3, generator=gen) * 0.3 + torch.tensor([0.0, 0.0, 1.0])
    d[:10] = -d[:10]                                            # rays looking away from the box
    d[10, 0] = 0.0                                              # a zero component (render_utils_kernel.cu:19: replaced by 1e-6)
    lo, hi = torch.tensor([-0.5, -0.4, -0.3]), torch.tensor([0.5, 0.6, 0.3])
    stepdist = 0.013
    pts, mask_out, ray_id, step_id, n_steps, t_min, t_max = dvgo_ops.sample_pts_on_rays(o, d, lo, hi, 2.0, 6.0, stepdist)
    assert bool((n_steps >= 1).all()) and int(n_steps.sum()) == len(pts) == len(ray_id) == len(step_id)
    assert bool((ray_id[1:] >= ray_id[:-1]).all())                                           # ray-major
    first = torch.cumsum(n_steps, 0) - n_steps
    assert torch.equal(step_id, torch.arange(len(pts)) - first[ray_id])                      # 0, 1, 2, ... inside every ray
    assert bool((t_min >= 2.0).all() and (t_max <= 6.0).all() and (t_min <= 6.0).all() and (t_max >= 2.0).all())
    miss = t_max < t_min                                        # the slabs do not overlap: the ray misses the box ...
    assert int(miss.sum()) > 0 and bool((n_steps[miss] == 1).all())          # ... and keeps its one obligatory sample,
    assert bool(mask_out[first[miss]].all())                                 # which lies outside
    assert torch.equal(n_steps, torch.clamp_min(torch.ceil((t_max - t_min) / np.float32(stepdist)), 1).long())
    inside = ~mask_out
    assert bool(((pts[inside] >= lo) & (pts[inside] <= hi)).all())
    assert int(inside[ray_id < 10].sum()) == 0                                               # looking away: nothing kept
    assert int(inside.sum()) > 1000
    # consecutive samples of a ray are stepdist apart along the unit direction
    same = ray_id[1:] == ray_id[:-1]
    gap = (pts[1:] - pts[:-1]).norm(dim=-1)[same]
    assert float((gap - stepdist).abs().max()) < 2e-6
    # no rays at all
    e = dvgo_ops.sample_pts_on_rays(o[:0], d[:0], lo, hi, 2.0, 6.0, stepdist)
    assert len(e[0]) == 0 and len(e[4]) == 0


def test_knn_contract_order_ties_and_float64_agreement():
    gen = torch.Generator().manual_seed(4)
    pts = torch.rand(500, 3, generator=gen)
    q = torch.rand(64, 3, generator=gen)
    d2, idx = knn_bruteforce(q, pts, 8)
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())                                             # ascending
    # the same neighbour SETS as a float64 search wherever the 8th and 9th distances are not within fp32 rounding of each other
    D = ((q.double()[:, None] - pts.double()[None]) ** 2).sum(-1)
    order = D.argsort(dim=1)
    clear = (D.gather(1, order[:, 8:9]) - D.gather(1, order[:, 7:8])).squeeze(1) > 1e-6
    assert int(clear.sum()) > 50
    assert torch.equal(idx[clear].sort(dim=1)[0], order[clear, :8].sort(dim=1)[0])
    # the distance itself: (dx*dx + dy*dy) + dz*dz in fp32 without FMA
    dx, dy, dz = [(q[:, None, i] - pts[None, :, i]) for i in range(3)]
    manual = ((dx * dx + dy * dy) + dz * dz).gather(1, idx)
    assert torch.equal(manual, d2)
    # exact ties: duplicated points come out in ascending index order
    dup = torch.cat([pts[:50], pts[:50], pts[:50]])                                           # every point three times
    d2t, it = knn_bruteforce(pts[:7] + 0.0, dup, 6)
    for r in range(7):
        for a, b in zip(range(5), range(1, 6)):
            assert (float(d2t[r, a]), int(it[r, a])) < (float(d2t[r, b]), int(it[r, b]))
        assert [int(v) for v in it[r, :3]] == [r, r + 50, r + 100] and float(d2t[r, 2]) == 0.0
    # K larger than the cloud is the caller's error in the reference (KeOps raises); the oracle needs N >= K
    assert knn_bruteforce(q[:2], pts[:8], 8)[1].sort(dim=1)[0].tolist() == [list(range(8))] * 2


def test_adam_step_size_is_the_float_expression_of_the_kernel():
    """lib/cuda/adam_upd_kernel.cu:72: step_size = lr * sqrt(1 - beta2^t) / (1 - beta1^t), evaluated in float."""
    for t in (1, 2, 10, 1000, 20000):
        ss = dvgo_ops._step_size(t, 0.9, 0.99, 1e-3)
        ref = 1e-3 * math.sqrt(1 - 0.99 ** t) / (1 - 0.9 ** t)
        assert abs(float(ss) - ref) <= 2e-6 * ref            # float evaluation: a few ulps off the double value
