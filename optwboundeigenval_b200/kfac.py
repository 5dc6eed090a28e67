"""K-FAC-preconditioned eigen-iteration (the reference's ``lobpcg=True`` variant).

Mirrors ``OptWBoundEignVal.init_kfac`` / ``kfac`` / the lobpcg branch of ``comp_rho``
(opt.py:362-416, 426-430, 491-493) and the pieces of ``kfac.py`` they use
(``_save_input``, ``_save_grad_output``, ``_update_inv``, ``_get_natural_grad``):

* factors ``A = 0.95 I + 0.05 cov(a)``, ``G = 0.95 I + 0.05 cov(g)`` per Conv2d/Linear module from ONE
  forward/backward over the batch -- built by ``b2s_kfac_build`` from the plan's cached base pass;
* ``eigh`` of both (kfac.py:87-93) stays a library call (``torch.linalg.eigh`` on the device), the
  inverses ``Q diag(1/d) Q^T`` are installed with ``b2s_kfac_set``;
* ``T r`` (``b2s_kfac_apply``) and ``v <- normalise(v + alpha(i) T r)`` run inside
  ``b2s_power_iterate(cfg.precond = 1)``.

``kfac_rand=True`` (targets sampled from the model output, opt.py:351-356) draws the sample on the host
with torch's CPU generator like the reference does.
"""
from __future__ import annotations

import ctypes
import time

import numpy as np
import torch

from . import _lib
from .hvp_operator import B200HVPOperator, _ptr
from .tracer import HEAD_CE, HEAD_SIGMOID_WBCE, HEAD_SOFTMAX_CE, HEAD_WBCE, OP_CONV


def _module_uses(tape):
    """module id -> (op index of its first forward use, op index of its last forward use)"""
    uses = {}
    for i, op in enumerate(tape.ops):
        if op.kind == OP_CONV and op.module is not None:
            first, _ = uses.get(id(op.module), (i, i))
            uses[id(op.module)] = (first, i)
    return list(uses.values())


def _sample_targets(op, loss_name):
    """opt.py:351-356: targets drawn from the model's own output distribution (CPU generator)."""
    plan = op.plan
    tape = plan.tape
    vt = tape.tensors[tape.logits]
    batch = op.size
    z = np.zeros((batch,) + vt.shape, dtype=np.float32)
    _lib.check(plan.lib.b2s_debug_read(plan.handle, 0, 0, tape.logits, z.ctypes.data_as(ctypes.c_void_p)))
    z = torch.from_numpy(z).view(batch, -1)
    if tape.head == HEAD_SOFTMAX_CE:
        out = torch.softmax(z, dim=1)
    elif tape.head == HEAD_SIGMOID_WBCE:
        out = torch.sigmoid(z)
    else:
        out = z
    if loss_name in ("W_BCEWithLogitsLoss", "WeightedBCEWithLogits", "BCELoss"):
        return torch.bernoulli(out).squeeze()
    return torch.multinomial(torch.softmax(out, dim=1), 1).squeeze()


def init_kfac(self, data):
    """opt.py:362-382."""
    start = time.time()
    op = B200HVPOperator(self.model, data, self.loss)
    op.prepare_grad()                                  # forward + backward over the batch (train mode)
    if getattr(self, "kfac_rand", False):
        inputs, _ = op.prep_data(data)
        sampled = _sample_targets(op, self.loss.__class__.__name__)
        op = B200HVPOperator(self.model, [inputs, sampled], self.loss)
        op.prepare_grad()
    plan = op.plan
    plan._bind_stream()
    keep = []
    _lib.check(plan.lib.b2s_kfac_clear(plan.handle))
    for first, last in _module_uses(plan.tape):
        da, dg = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(plan.lib.b2s_kfac_dims(plan.handle, last, ctypes.byref(da), ctypes.byref(dg)))
        A = torch.empty(da.value, da.value, dtype=torch.float32, device=plan.device)
        G = torch.empty(dg.value, dg.value, dtype=torch.float32, device=plan.device)
        # the reference's hooks leave A from the LAST forward use and G from the FIRST (backward visits
        # uses in reverse) when a module is applied twice (forest_data.py:85-86)
        _lib.check(plan.lib.b2s_kfac_build(plan.handle, last, first, _ptr(A), _ptr(G)), "b2s_kfac_build")
        inv = []
        for M in (A, G):                               # kfac.py:87-93, 118-120
            d, Q = torch.linalg.eigh(M)
            d = d * (d > 1e-10).float()
            inv.append(((Q / d.unsqueeze(0)) @ Q.t()).contiguous())
        _lib.check(plan.lib.b2s_kfac_set(plan.handle, first, _ptr(inv[0]), _ptr(inv[1])), "b2s_kfac_set")
        keep.extend(inv)
    plan._kfac_keep = keep                             # the library borrows these buffers
    self.kTime = getattr(self, "kTime", 0) + time.time() - start


def kfac(self, r):
    """opt.py:384-416: T r on an fp64 device vector."""
    start = time.time()
    plan = self.hvp_op.plan if getattr(self, "hvp_op", None) is not None and self.hvp_op.plan is not None else None
    if plan is None:
        raise RuntimeError("kfac(): no cached plan; call comp_rho first")
    r = r.to(plan.device, torch.float64).contiguous()
    out = torch.empty_like(r)
    plan._bind_stream()
    _lib.check(plan.lib.b2s_kfac_apply(plan.handle, _ptr(r), _ptr(out)), "b2s_kfac_apply")
    self.kTime = getattr(self, "kTime", 0) + time.time() - start
    return out


def preconditioned_comp_rho(self, data, p=False):
    """lobpcg branch of comp_rho (opt.py:426-430, 447-520); self.hvp_op was just created by comp_rho."""
    from .spectral import _finish_rho
    if self.kfac_iter >= self.kfac_batch:
        init_kfac(self, data)
        self.kfac_iter = 1
    else:
        self.kfac_iter += 1
    v = self.random_v() if self.rand_init else self.v
    n_steps = int(np.min([self.ndim, self.max_pow_iter]))
    alpha = self.pow_iter_alpha
    alphas = [float(alpha(i)) if callable(alpha) else float(alpha) for i in range(n_steps)]
    pstart = time.time()
    self.hvp_op._ensure_grad(True)
    out = self.hvp_op.plan.power_iterate(self.hvp_op._vec(v), self.pow_iter_eps, n_steps, alphas,
                                         want_trajectory=bool(self.verbose), precond=True)
    p_time = time.time() - pstart
    _finish_rho(self, out, p, p_time, p_time)
    return out.iters, out.rn, self.hvp_op.size
