"""CPU oracle for the spectral-radius hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, on plain CPU PyTorch autograd, the algorithm the reference
runs for the hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does (it fails loudly when its CUDA library is missing).

Where the arithmetic lives: third-party **PyTorch autograd** (not vendored by the
reference, no version pin there; this image has torch 2.11.0).  The reference's
call sites are ``opt.py:99,132,143,189``; each function below cites the lines it
follows.

Pinning: ``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (in the build container, where it exists) and stores its
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function of this file against those vectors.  The reference itself ships no
golden vectors for this path.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch


def flat_params(model) -> torch.Tensor:
    """Parameter vector in ``model.parameters()`` order (opt.py:102,191)."""
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def _flatten(grads) -> torch.Tensor:
    return torch.cat([g.contiguous().view(-1) for g in grads])


class AutogradSpectralOperator:
    """gradient / Hv / vGHv by nested reverse mode.

    Follows ``HVPOperator`` (opt.py:48-192): the gradient is built once with
    ``create_graph=True`` and kept; ``hv`` differentiates ``<grad, vec>`` again;
    ``vghv`` differentiates ``<Hv, vec>`` a third time.  Unlike the reference
    object this one keeps the graph alive, so ``vghv`` may be called repeatedly.
    """

    def __init__(self, model, data, criterion):
        self.model = model
        self.criterion = criterion
        if isinstance(data, dict):          # chest datasets (opt.py:168-169)
            self.inputs, self.target = data["image"], data["label"]
        else:                               # [inputs, target] (opt.py:164-167)
            self.inputs, self.target = data
        self.size = len(self.target)
        self._g = None
        self.loss_value = None

    def params(self):
        return list(self.model.parameters())

    def gradient(self) -> torch.Tensor:
        """opt.py:175-192 (prepare_grad): forward, loss, grad with graph, flatten, cast to fp64."""
        if self._g is None:
            out = self.model(self.inputs)
            if self.criterion.__class__.__name__ == "KLDivLoss":       # opt.py:182-185: one-hot target
                onehot = torch.zeros(out.shape)
                onehot.scatter_(1, self.target.view(-1, 1), 1)
                loss = self.criterion(out.float(), onehot.float())
            else:
                loss = self.criterion(out, self.target)
            self.loss_value = float(loss.detach())
            g = torch.autograd.grad(loss, self.params(), create_graph=True)
            self._g = _flatten(g).double()
        return self._g

    def hv(self, vec) -> torch.Tensor:
        """opt.py:77-108."""
        vec = torch.as_tensor(vec).double()
        gg = torch.autograd.grad(self.gradient(), self.params(), grad_outputs=vec, retain_graph=True)
        return _flatten(gg).detach().double()

    def vghv(self, vec) -> torch.Tensor:
        """opt.py:110-152: grad of <H vec, vec> with vec held fixed."""
        vec = torch.as_tensor(vec).double()
        hv = torch.autograd.grad(self.gradient(), self.params(), grad_outputs=vec, create_graph=True)
        hv = _flatten(hv).double()
        ggg = torch.autograd.grad(hv, self.params(), grad_outputs=vec, retain_graph=True)
        return _flatten(ggg).detach().double()


def start_vector(ndim: int) -> torch.Tensor:
    """opt.py:324-325: the 'random' start vector is ones/sqrt(P)."""
    return torch.from_numpy(np.ones(ndim) / np.sqrt(ndim))


def power_iteration(hv: Callable[[torch.Tensor], torch.Tensor], v0: torch.Tensor, *, eps: float,
                    max_iter: int, alpha=1.0, precond: Optional[Callable] = None):
    """Spectral-radius iteration of ``comp_rho`` (opt.py:447-520).

    Returns a dict with the final ``v`` (the vector the last Hv was taken at when a
    stopping test fired; the updated one when iterations ran out), ``lam`` (signed
    flip applied, so >= 0), ``norm`` (residual), ``rn``, ``iters`` (index of the last
    iteration, as the reference returns ``i``), ``converged`` and the per-iteration
    ``trajectory`` rows ``(i, lam, n, rn, vnn)`` that the verbose log prints
    (opt.py:466).
    """
    v = v0.clone().double()
    ndim = v.numel()
    n_steps = int(min(ndim, max_iter))
    lam = n = 0.0
    r_old = 0.0
    n_old = lam_old = 0.0
    rn = 0.0
    stop = [math.inf] * 3
    traj = []
    i = -1
    for i in range(n_steps):
        w = hv(v)
        lam = float(torch.dot(w, v))
        if lam < 0:                       # opt.py:458-460
            lam, w = -lam, -w
        r = w - lam * v
        n = float(torch.norm(r))
        rn = float(min(torch.norm(r - r_old), torch.norm(r + r_old)))   # opt.py:463
        vnn = float(torch.norm(w))
        traj.append((i, lam, n, rn, vnn))
        stop = [n,
                rn / n_old if n_old != 0 else math.inf,
                abs(lam - lam_old) / lam_old if lam_old != 0 else math.inf]   # opt.py:479
        if any(s < eps for s in stop):
            break
        if i < n_steps - 1:               # opt.py:483-485 (`reset` is never set)
            lam_old, r_old, n_old = lam, r, n
        a = alpha(i) if callable(alpha) else alpha
        if precond is not None:           # opt.py:491-493
            w = v + a * precond(r)
        else:                             # opt.py:495
            w = v + a * (w - v)
        v = w / torch.norm(w)             # opt.py:498
    converged = not all(s > eps for s in stop)      # opt.py:513
    return {"v": v, "lam": lam, "rho": abs(lam), "norm": n, "rn": rn, "iters": i,
            "converged": converged, "trajectory": traj, "stop": stop}


def penalty_gradient(op: AutogradSpectralOperator, v: torch.Tensor, clip: Optional[float] = None):
    """comp_gradrho (opt.py:535-542): vGHv at v, optionally rescaled to norm ``clip``."""
    g = op.vghv(v)
    if clip is not None:
        gn = float(torch.norm(g))
        if gn > clip:
            g = g * (clip / gn)
    return g


def regularizer_value(rho: float, K: float, Kmin: float = 0.0) -> float:
    """comp_g (opt.py:578)."""
    return max(0.0, rho - K, Kmin - rho)


def step_direction(gradf: torch.Tensor, gradrho: Optional[torch.Tensor], rho: float, K: float, mu: float,
                   g: float) -> torch.Tensor:
    """iter() assembly (opt.py:631-639): grad f + mu * sign * grad rho when the penalty is active."""
    if g > 0 and gradrho is not None:
        sign = 1.0 if rho > K else -1.0
        return gradf + mu * sign * gradrho
    return gradf.clone()


# --------------------------------------------------------------------------------------
# K-FAC preconditioner used when lobpcg=True (opt.py:362-416 with kfac.py:50-130,277-367)
# --------------------------------------------------------------------------------------

def _patches(x: torch.Tensor, conv) -> torch.Tensor:
    """kfac.py:201-218: (B, oh, ow, cin*kh*kw) patch matrix of a conv input."""
    ph, pw = conv.padding
    if ph + pw > 0:
        x = torch.nn.functional.pad(x, (pw, pw, ph, ph))
    kh, kw = conv.kernel_size
    sh, sw = conv.stride
    x = x.unfold(2, kh, sh).unfold(3, kw, sw)           # B, C, oh, ow, kh, kw
    x = x.permute(0, 2, 3, 1, 4, 5).contiguous()
    return x.view(x.size(0), x.size(1), x.size(2), -1)


def kfac_cov_a(a: torch.Tensor, layer) -> torch.Tensor:
    """kfac.py:292-311: activation second moment, bias column of ones appended."""
    batch = a.size(0)
    if isinstance(layer, torch.nn.Conv2d):
        a = _patches(a, layer)
        spatial = a.size(1) * a.size(2)
        a = a.view(-1, a.size(-1))
        if layer.bias is not None:
            a = torch.cat([a, a.new_ones(a.size(0), 1)], 1)
        a = a / spatial
    else:
        if layer.bias is not None:
            a = torch.cat([a, a.new_ones(a.size(0), 1)], 1)
    return a.t() @ (a / batch)


def kfac_cov_g(g: torch.Tensor, layer, batch_averaged: bool = True) -> torch.Tensor:
    """kfac.py:337-367: output-gradient second moment."""
    batch = g.size(0)
    if isinstance(layer, torch.nn.Conv2d):
        spatial = g.size(2) * g.size(3)
        g = g.permute(0, 2, 3, 1).contiguous().view(-1, g.size(1))
        if batch_averaged:
            g = g * batch
        g = g * spatial
        return g.t() @ (g / g.size(0))
    if batch_averaged:
        return g.t() @ (g * batch)
    return g.t() @ (g / batch)


class KfacPreconditioner:
    """Factors on one batch and the layer-wise map ``r -> G^-1 R A^-1``.

    Follows init_kfac (opt.py:362-382): one forward/backward over the batch with the
    TRUE targets (``kfac_rand=False`` path, opt.py:357-358), factor buffers
    ``0.95*I + 0.05*cov`` (kfac.py:52-65 with ``steps == 0`` always), ``eigh`` with
    eigenvalues below 1e-10 zeroed (kfac.py:87-93); and kfac() (opt.py:384-416) with
    damping 0 (kfac.py:118-120).  For a layer used more than once in the forward
    (the forest MLP's shared fc2) the LAST hook to fire wins: the last use in the
    forward for A, the first use for G (backward visits uses in reverse).
    """

    def __init__(self, model, stat_decay: float = 0.95):
        self.model = model
        self.decay = stat_decay
        self.layers = [m for m in model.modules() if m.__class__.__name__ in ("Linear", "Conv2d")]
        self.Qa, self.Qg, self.da, self.dg = {}, {}, {}, {}

    def build(self, data, criterion):
        inputs, target = (data["image"], data["label"]) if isinstance(data, dict) else data
        acts, gouts, handles = {}, {}, []
        for m in self.layers:
            handles.append(m.register_forward_pre_hook(lambda mod, inp: acts.__setitem__(mod, inp[0].detach())))
            handles.append(m.register_full_backward_hook(
                lambda mod, gi, go: gouts.__setitem__(mod, go[0].detach())))
        try:
            x = inputs.clone().requires_grad_()
            out = self.model(x)
            loss = criterion(out, target)
            loss.backward()
        finally:
            for h in handles:
                h.remove()
        for m in self.layers:
            aa = kfac_cov_a(acts[m], m)
            gg = kfac_cov_g(gouts[m], m, True)
            maa = self.decay * torch.eye(aa.size(0)) + (1 - self.decay) * aa
            mgg = self.decay * torch.eye(gg.size(0)) + (1 - self.decay) * gg
            self.da[m], self.Qa[m] = torch.linalg.eigh(maa)
            self.dg[m], self.Qg[m] = torch.linalg.eigh(mgg)
            self.da[m] = self.da[m] * (self.da[m] > 1e-10).float()
            self.dg[m] = self.dg[m] * (self.dg[m] > 1e-10).float()
        self.model.zero_grad()

    def apply(self, r: torch.Tensor) -> torch.Tensor:
        """opt.py:384-416: identity on parameters of modules that are not Linear/Conv2d."""
        out = r.clone()
        j = 0
        for m in self.model.modules():
            own = list(m.parameters(recurse=True))
            n_own = len(own)
            is_leaf = m.__class__.__name__ != "Sequential" and (
                (n_own == 2 and getattr(m, "bias", None) is not None)
                or (n_own == 1 and hasattr(m, "bias") and m.bias is None)
                or (n_own == 1 and not hasattr(m, "bias")))
            if not is_leaf:
                continue
            sizes = [p.numel() for p in own]
            total = sum(sizes)
            if m in self.Qa:
                mat = r[j:j + sizes[0]].view(own[0].shape[0], -1).float()
                if getattr(m, "bias", None) is not None:
                    mat = torch.cat([mat, r[j + sizes[0]:j + total].view(-1, 1).float()], 1)
                core = self.Qg[m].t() @ mat @ self.Qa[m]
                core = core / (self.dg[m].unsqueeze(1) * self.da[m].unsqueeze(0) + 0)
                nat = self.Qg[m] @ core @ self.Qa[m].t()
                if getattr(m, "bias", None) is not None:
                    flat = torch.cat([nat[:, :-1].reshape(-1), nat[:, -1:].reshape(-1)])
                else:
                    flat = nat.reshape(-1)
                out[j:j + total] = flat.to(out.dtype)
            j += total
        return out
