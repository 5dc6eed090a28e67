"""Fused replacements for ``OptWBoundEignVal.comp_rho / comp_gradrho / comp_g``
(reference opt.py:418-542, 574-578).

``comp_rho`` in the reference is a Python loop of ~15 ATen vector kernels and
>= 5 host syncs per iteration around ``HVPOperator.Hv``.  Here the whole loop
runs in ``b2s_power_iterate``: one captured HVP graph plus two fused vector
kernels per iteration, scalars resident on the device.  Side effects are kept
bit-for-bit in *kind*: ``self.hvp_op``, ``self.v`` (fp64 device tensor),
``self.rho`` (float, -1 sentinel), ``self.norm``, the verbose log lines
(opt.py:443,466,503-504), the warnings, and the return value
``(i, rn, batch_size)``.

These functions take the trainer as ``self`` so they can be bound onto the
reference class (``dropin.install``) or used with ``SpectralState`` below.
"""
from __future__ import annotations

import sys
import time

import numpy as np
import torch

from .hvp_operator import B200HVPOperator


def _time_hms(t, head=""):
    """same text as opt.timeHMS (opt.py:230-235)"""
    hrs = np.floor(t / 3600)
    t = t - hrs * 3600
    mins = np.floor(t / 60)
    secs = t - mins * 60
    print(head + "Time elapsed: %2i hrs, %2i min, %4.2f sec" % (hrs, mins, secs))


def comp_rho(self, data, p=False):
    """opt.py:418-533."""
    self.model.train()
    self.hvp_op = B200HVPOperator(self.model, data, self.loss, use_gpu=True)
    lobpcg = bool(getattr(self, "lobpcg", False))
    if lobpcg:
        from .kfac import preconditioned_comp_rho
        return preconditioned_comp_rho(self, data, p)

    v = self.random_v() if self.rand_init else self.v
    n_steps = int(np.min([self.ndim, self.max_pow_iter]))
    alpha = self.pow_iter_alpha
    if callable(alpha):
        alphas = [float(alpha(i)) for i in range(n_steps)]
    elif float(alpha) != 1.0:
        alphas = [float(alpha)] * n_steps
    else:
        alphas = None

    pstart = time.time()
    out = self.hvp_op.power_iterate(v, self.pow_iter_eps, n_steps, alphas, want_trajectory=bool(self.verbose))
    p_time = time.time() - pstart
    _finish_rho(self, out, p, p_time, p_time)
    return out.iters, out.rn, self.hvp_op.size


def _finish_rho(self, out, p, hv_time, p_time):
    if self.verbose:
        old_stdout = sys.stdout
        with open(self.verbose_log_file, "a") as log_file:
            sys.stdout = log_file
            try:
                print("iter\t lam\t norm\t delRes\t vnnorm")
                for row in out.trajectory:
                    print("%d\t %f\t %f\t %f\t %f" % (int(row[0]), row[1], row[2], row[3], row[4]))
                _time_hms(hv_time, "HV ")
                _time_hms(p_time, "Power Iter ")
            finally:
                sys.stdout = old_stdout

    self.v = out.v
    self.rho = np.abs(out.lam)
    self.norm = out.norm
    if all(s > self.pow_iter_eps for s in out.stop):          # opt.py:513-520
        pr = "Warning: power iteration has not fully converged."
        if self.ignore_bad_vals:
            pr += " Ignoring rho."
            self.rho = -1
            self.v = self.random_v()
        print(pr)
    if out.lam == 0:
        print("Warning: rho = 0")
    if p:
        old_stdout = sys.stdout
        with open(self.log_file, "a") as log_file:
            sys.stdout = log_file
            try:
                print("Rho:", self.rho)
            finally:
                sys.stdout = old_stdout


def comp_gradrho(self):
    """opt.py:535-542."""
    self.gradrho = self.hvp_op.vGHv(self.v, storedGrad=True)
    if self.gradg_clip is not None:
        grn = torch.norm(self.gradrho)
        if grn > self.gradg_clip:
            self.gradrho *= self.gradg_clip / grn


def comp_g(self, data):
    """opt.py:574-578."""
    self.comp_rho(data)
    self.g = np.max([0.0, self.rho - self.K, self.Kmin - self.rho])


def rho_test(self, loader, replicas=True):
    """``OptWBoundEignVal.rho_test`` (opt.py:882-910): lambda_max of every minibatch of ``loader`` and the batch-size
    weighted averages of ``[rho, norm, iterations, rn, seconds]``.

    With ``torch.distributed`` initialised and ``replicas=True`` the sweep is *replicas only*: rank r takes the
    minibatches j with j % world == r, runs the single-GPU path on them (no synced BatchNorm, no all-reduce: every
    minibatch is a whole batch on one GPU, exactly as in the reference) and the per-batch rows are gathered at the
    end -- the path has no exchange step, so none is invented.  Returns ``(stats, averages)``; every rank gets both."""
    import torch.distributed as dist
    from . import hvp_operator
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    sharded = world > 1 and replicas
    if sharded:
        hvp_operator.set_data_parallel(False)
    try:
        rows, sizes = [], []
        for j, data in enumerate(loader):
            if sharded and j % world != rank:
                continue
            start = time.time()
            i, rn, size = self.comp_rho(data)
            rows.append([j, float(self.rho), float(self.norm), int(i), float(rn), time.time() - start])
            sizes.append(int(size))
    finally:
        if sharded:
            hvp_operator.set_data_parallel(True)
    if sharded:
        gathered = [None] * world
        dist.all_gather_object(gathered, (rows, sizes))
        rows = [r for part in gathered for r in part[0]]
        sizes = [z for part in gathered for z in part[1]]
        order = np.argsort([r[0] for r in rows], kind="stable")
        rows = [rows[k] for k in order]
        sizes = [sizes[k] for k in order]
    stats = np.array(rows, dtype="float")
    avg = np.average(stats, axis=0, weights=sizes)[1:] if len(rows) else np.zeros(5)
    return stats, avg


class _StepBuffers(object):
    """Flat fp64 / fp32 step vectors of one model and the per-parameter views into the fp32 one (computed once:
    the reference recomputes ``torch.prod(torch.tensor(s))`` and slices + casts per parameter on every minibatch,
    opt.py:654-659)."""

    def __init__(self, model, device):
        self.params = list(model.parameters())
        self.n = sum(q.numel() for q in self.params)
        self.p64 = torch.empty(self.n, dtype=torch.float64, device=device)
        self.p32 = torch.empty(self.n, dtype=torch.float32, device=device)
        self.views = []
        i = 0
        for q in self.params:
            self.views.append(self.p32[i:i + q.numel()].view(q.size()))
            i += q.numel()


def assemble_step(self, mu=None):
    """The step assembly of iter() (opt.py:622-639, 654-659) after ``comp_g``: ``p = grad f + mu * sign * grad rho`` in one
    fused pass (C ABI ``b2s_step_assemble``) and ``param.grad`` set to views of the flat fp32 copy.  Returns ``p`` (fp64)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    buf = getattr(self, "_step_buffers", None)
    if buf is None or buf.params != list(self.model.parameters()):
        buf = self._step_buffers = _StepBuffers(self.model, self.device)
    if mu is None:
        mu = self.mu(getattr(self, "i", 0)) if callable(self.mu) else self.mu
    if self.hvp_op.stored_grad is not None:                                   # opt.py:624-627
        self.gradf = self.hvp_op.stored_grad.data.to(self.device)
    else:
        self.gradf = torch.zeros(self.ndim, dtype=torch.float64, device=self.device)
    coef, gr = 0.0, None
    if self.g > 0:                                                            # opt.py:631-637
        self.comp_gradrho()
        sign = 1 if self.rho > self.K else -1
        self.gradg = sign * self.gradrho if getattr(self, "keep_gradg", False) else None
        coef, gr = float(mu) * sign, self.gradrho
    st = torch.cuda.current_stream(self.device).cuda_stream
    _lib.check(lib.b2s_step_assemble(ctypes.c_void_p(self.gradf.data_ptr()), ctypes.c_void_p(gr.data_ptr()) if gr is not None else None,
                                     ctypes.c_double(coef), buf.n, ctypes.c_void_p(buf.p64.data_ptr()),
                                     ctypes.c_void_p(buf.p32.data_ptr()), ctypes.c_void_p(st)), "b2s_step_assemble")
    for q, view in zip(buf.params, buf.views):                               # opt.py:654-659 without the slicing / casting kernels
        q.grad = view
    return buf.p64


class SpectralState(object):
    """Minimal stand-alone carrier of the attributes the three functions use -- what
    ``OptWBoundEignVal.__init__`` (opt.py:239-316) sets up -- for users (bench, tests, smoke) that
    do not have the reference on their path."""

    def __init__(self, model, loss, mu=0, K=0, Kmin=0, pow_iter_eps=1e-3, max_pow_iter=1000, pow_iter_alpha=1,
                 rand_init=False, ignore_bad_vals=True, gradg_clip=None, verbose=False, lobpcg=False, kfac_batch=1,
                 kfac_rand=True, verbose_log_file="./logs/spectral_verbose.log", log_file="./logs/spectral.log"):
        if not torch.cuda.is_available():
            raise RuntimeError("optwboundeigenval_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.model = model.to(self.device)
        self.loss = loss
        self.ndim = sum(p.numel() for p in model.parameters())
        self.mu, self.K, self.Kmin = mu, float(K), float(Kmin)
        self.pow_iter_eps, self.max_pow_iter, self.pow_iter_alpha = pow_iter_eps, max_pow_iter, pow_iter_alpha
        self.rand_init, self.ignore_bad_vals, self.gradg_clip = rand_init, ignore_bad_vals, gradg_clip
        self.verbose, self.verbose_log_file, self.log_file = verbose, verbose_log_file, log_file
        self.lobpcg, self.kfac_batch, self.kfac_iter, self.kfac_rand = lobpcg, kfac_batch, kfac_batch, kfac_rand
        self.use_gpu = True
        self.hvp_op = None
        self.rho = 0
        self.norm = 0
        self.g = 0
        self.v = self.random_v()
        self.gradrho = torch.zeros(self.ndim, dtype=torch.float64, device=self.device)
        self.kTime = 0

    def random_v(self):   # opt.py:324-325
        return torch.from_numpy(1.0 / np.sqrt(self.ndim) * np.ones(self.ndim)).to(self.device)

    comp_rho = comp_rho
    comp_gradrho = comp_gradrho
    comp_g = comp_g

    def kfac(self, r):            # opt.py:384-416
        from .kfac import kfac as _kfac
        return _kfac(self, r)

    def init_kfac(self, data):    # opt.py:362-382
        from .kfac import init_kfac as _init
        return _init(self, data)

    assemble_step = assemble_step
    rho_test = rho_test

    def regularized_step(self, data, optimizer):
        """One minibatch of iter() (opt.py:608-699, the pow_iter branch with a plain optimizer): comp_g, grad f,
        penalty gradient when g > 0, fused step assembly into param.grad, optimizer.step()."""
        self.comp_g(data)
        optimizer.zero_grad()
        p = self.assemble_step()
        optimizer.step()
        return p

    def step_direction(self, data):
        """The assembly of iter() (opt.py:616-639): grad f + mu * sign * grad rho."""
        self.comp_g(data)
        gradf = self.hvp_op.stored_grad.data.to(self.device)
        mu = self.mu(0) if callable(self.mu) else self.mu
        if self.g > 0:
            self.comp_gradrho()
            sign = 1 if self.rho > self.K else -1
            gradg = sign * self.gradrho
        else:
            gradg = torch.zeros(self.ndim, dtype=torch.float64, device=self.device)
        return gradf + mu * gradg
