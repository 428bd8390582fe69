"""MaskedAdam with the reference's interface (lib/masked_adam.py:17-72), served by the multi-tensor
kernel in csrc/adam.cu: one launch updates every parameter tensor of every group instead of one
launch per tensor.

Update rule (lib/cuda/adam_upd_kernel.cu:9-58,72): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
p -= lr sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps); `skip_zero_grad` groups skip elements with g == 0.
"""
from __future__ import annotations

import torch

from . import ops


class MaskedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        defaults = dict(lr=lr, betas=betas, eps=eps, skip_zero_grad=False)
        self.per_lr = None
        super().__init__(params, defaults)

    def set_pervoxel_lr(self, count):
        assert self.param_groups[0]['params'][0].shape == count.shape
        self.per_lr = (count.float() / count.max()).contiguous()

    @torch.no_grad()
    def prepare_step(self):
        """Host half of a step (no launch): advances the step counters, (re)builds the cached descriptor tables and returns
        [(class (beta1, beta2, eps), AdamPlan, [step size per tensor]), ...] plus the parameters that will be updated.
        `step()` launches them right away; train.GraphedTrainStep writes the step sizes to device memory and replays a
        captured graph that holds the launches."""
        plan = []
        for group in self.param_groups:
            lr = group['lr']
            beta1, beta2 = group['betas']
            eps = group['eps']
            skip_zero_grad = group.get('skip_zero_grad', False)
            ss_cache = {}
            for param in group['params']:
                if param.grad is None:
                    continue
                state = self.state[param]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(param, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(param, memory_format=torch.preserve_format)
                state['step'] += 1
                grad = param.grad if param.grad.is_contiguous() else param.grad.contiguous()
                if self.per_lr is not None and param.shape == self.per_lr.shape:
                    mode, perlr = 2, self.per_lr
                elif skip_zero_grad:
                    mode, perlr = 1, None
                else:
                    mode, perlr = 0, None
                st = state['step']
                ss = ss_cache.get(st)
                if ss is None:
                    ss = ss_cache[st] = ops.adam_step_size(st, beta1, beta2, lr)
                plan.append(((beta1, beta2, eps), param, grad, state, perlr, ss, mode))
        if not plan:
            return [], []
        # every pointer the descriptor table holds is part of the key: load_state_dict() replaces the moment tensors
        # (the reference reads self.state[param] afresh every step, lib/masked_adam.py:52-60)
        key = tuple((cls, p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), mode,
                     None if pl is None else pl.data_ptr()) for cls, p, g, st, pl, _, mode in plan)
        if getattr(self, '_plan_key', None) != key:
            batches = {}
            for cls, p, g, state, pl, ss, mode in plan:
                batches.setdefault(cls, []).append((p.data, g, state['exp_avg'], state['exp_avg_sq'], pl, ss, mode))
            self._plan = [(cls, ops.AdamPlan(entries)) for cls, entries in batches.items()]
            self._plan_key = key
        return ([(cls, ap, [e[5] for e in plan if e[0] == cls]) for cls, ap in self._plan], [e[1] for e in plan])

    def undo_step_count(self):
        """Takes back the step counters of the last prepare_step() (a step whose device-side update was skipped)."""
        for group in self.param_groups:
            for param in group['params']:
                st = self.state.get(param)
                if st and st.get('step', 0) > 0 and param.grad is not None:
                    st['step'] -= 1

    @torch.no_grad()
    def step(self):
        """One multi-tensor launch per (betas, eps) class.  The descriptor table (device pointers of param / grad /
        moments) is cached and rebuilt only when a pointer changes; per step only the step sizes are refreshed."""
        launches, params = self.prepare_step()
        for cls, ap, sizes in launches:
            ap.launch(sizes, *cls)
        # the kernel wrote through raw pointers: tell autograd's version counters, which key every derived cache
        # (packed tensor-core weights, per-point layer-0 table, ...) exactly as an in-place torch op would
        if params:
            torch.autograd.graph.increment_version(params)
