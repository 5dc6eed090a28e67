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


def comp_f(self, inputs, target, classes=None, model_classes=None):
    """``OptWBoundEignVal.comp_f`` (opt.py:544-572): loss and model output of one batch in evaluation mode, forward only
    (C ABI ``b2s_eval_pass``: evaluation-mode BatchNorm, no adjoint sweep).  ``test_model`` (opt.py:912-1039) makes all
    of its forward passes through this method, so it is covered as well.  With ``classes`` (a subset of the label
    columns) the loss is taken on the subset of the returned output exactly as the reference does."""
    from .hvp_operator import flat_parameters, plan_for
    self.model.eval()
    inputs = inputs.to(self.device)
    target = target.to(self.device)
    plan = plan_for(self.model, self.loss, inputs, self.device)
    with torch.no_grad():
        loss, output = plan.eval_pass(flat_parameters(self.model), inputs, target)
        if classes is not None:
            if model_classes is None:
                model_classes = classes
            if target.shape[1] == 1:
                print('"Classes" argument only implemented for one-hot encoding')
            else:
                return self.loss(output[:, model_classes], target[:, classes]).item(), output[:, model_classes]
    return float(loss.item()), output


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
        self.gradg = sign * self.gradrho
        coef, gr = float(mu) * sign, self.gradrho
    else:
        self.gradg = torch.zeros(self.ndim, dtype=torch.float64, device=self.device)      # opt.py:636
    st = torch.cuda.current_stream(self.device).cuda_stream
    _lib.check(lib.b2s_step_assemble(ctypes.c_void_p(self.gradf.data_ptr()), ctypes.c_void_p(gr.data_ptr()) if gr is not None else None,
                                     ctypes.c_double(coef), buf.n, ctypes.c_void_p(buf.p64.data_ptr()),
                                     ctypes.c_void_p(buf.p32.data_ptr()), ctypes.c_void_p(st)), "b2s_step_assemble")
    for q, view in zip(buf.params, buf.views):                               # opt.py:654-659 without the slicing / casting kernels
        q.grad = view
    return buf.p64


class _FusedOptimizer(object):
    """Flat state of a ``torch.optim.SGD`` / ``torch.optim.Adam`` whose parameters are views of one flat fp32 vector
    (``hvp_operator.FlatParams.attach``), so that the update of opt.py:696-699 runs inside ``b2s_step_fused``.

    The optimizer object stays the owner of its hyper-parameters (``param_groups`` are read every step, so LR
    schedulers keep working) and of its state: ``optimizer.state[p]`` holds VIEWS of the flat state vectors under
    torch's own keys (``momentum_buffer`` / ``step``, ``exp_avg``, ``exp_avg_sq``), so ``state_dict()``, checkpoints and
    a later plain ``optimizer.step()`` see exactly what torch would have produced."""

    def __init__(self, optimizer, flat):
        self.optimizer, self.flat = optimizer, flat
        self.kind = 1 if type(optimizer) is torch.optim.SGD else 2
        dev = flat.flat.device
        self.s1 = torch.zeros(flat.n, dtype=torch.float32, device=dev)
        self.s2 = torch.zeros(flat.n, dtype=torch.float32, device=dev) if self.kind == 2 else None
        index = {id(q): k for k, q in enumerate(flat.params)}
        self.groups = []                       # (group dict, [(offset, n)], [param indices])
        for g in optimizer.param_groups:
            ks = sorted(index[id(q)] for q in g["params"])
            segs = []
            for k in ks:
                o, n = flat.offsets[k], flat.sizes[k]
                if segs and segs[-1][0] + segs[-1][1] == o:
                    segs[-1][1] += n
                else:
                    segs.append([o, n])
            self.groups.append((g, segs, ks))
        self.bound = False

    @staticmethod
    def supported(optimizer, flat):
        if type(optimizer) not in (torch.optim.SGD, torch.optim.Adam):
            return False
        known = {id(q) for q in flat.params}
        for g in optimizer.param_groups:
            if any(id(q) not in known for q in g["params"]) or g.get("differentiable") or g.get("capturable"):
                return False
            if not isinstance(g["lr"], (int, float)):
                return False
            if type(optimizer) is torch.optim.Adam and (g.get("amsgrad") or g.get("decoupled_weight_decay")):
                return False
        return True

    def _view(self, buf, k):
        o, n = self.flat.offsets[k], self.flat.sizes[k]
        return buf[o:o + n].view(self.flat.params[k].shape)

    def _bind(self):
        """adopt whatever state the optimizer already has, then point its entries at the flat vectors"""
        st = self.optimizer.state
        for g, _, ks in self.groups:
            for k in ks:
                q = self.flat.params[k]
                if self.kind == 1:
                    e = st.get(q)                      # plain SGD keeps no state: do not create entries
                    if e is not None and e.get("momentum_buffer") is not None:
                        v = self._view(self.s1, k)
                        if e["momentum_buffer"].data_ptr() != v.data_ptr():
                            v.copy_(e["momentum_buffer"])
                        e["momentum_buffer"] = v
                else:
                    e = st[q]
                    if "step" not in e:
                        e["step"] = torch.tensor(0.0, dtype=torch.float32)
                    for key, buf in (("exp_avg", self.s1), ("exp_avg_sq", self.s2)):
                        v = self._view(buf, k)
                        if key in e and e[key].data_ptr() != v.data_ptr():
                            v.copy_(e[key])
                        e[key] = v
        self.bound = True

    def step(self, lib, gradf, gradrho, coef, scale2, p32, stream):
        import ctypes
        from . import _lib
        if not self.bound:
            self._bind()
        st = self.optimizer.state
        w = self.flat.flat
        for g, segs, ks in self.groups:
            o = _lib.StepOpt()
            o.kind = self.kind
            o.maximize = 1 if g.get("maximize") else 0
            o.write_gradrho = 0
            o.lr, o.weight_decay = float(g["lr"]), float(g["weight_decay"])
            if self.kind == 1:
                o.momentum, o.dampening, o.nesterov = float(g["momentum"]), float(g["dampening"]), 1 if g["nesterov"] else 0
                have = [(st.get(self.flat.params[k]) or {}).get("momentum_buffer") is not None for k in ks]
                if o.momentum != 0 and any(have) != all(have):
                    raise RuntimeError("fused SGD step: only some parameters of a group have a momentum buffer")
                o.first_step = 0 if (o.momentum == 0 or all(have)) else 1
            else:
                steps = [st[self.flat.params[k]]["step"] for k in ks]
                torch._foreach_add_(steps, 1)
                t = float(steps[0])
                b1, b2 = g["betas"]
                o.beta1, o.beta2, o.eps = float(b1), float(b2), float(g["eps"])
                o.step_size = o.lr / (1.0 - b1 ** t)
                o.bias2_sqrt = (1.0 - b2 ** t) ** 0.5
            for off, n in segs:
                _lib.check(lib.b2s_step_fused(
                    ctypes.c_void_p(gradf.data_ptr() + 8 * off),
                    ctypes.c_void_p(gradrho.data_ptr() + 8 * off) if gradrho is not None else None,
                    ctypes.c_double(coef), ctypes.c_void_p(scale2.data_ptr()) if scale2 is not None else None, n, None,
                    ctypes.c_void_p(p32.data_ptr() + 4 * off), ctypes.c_void_p(w.data_ptr() + 4 * off),
                    ctypes.c_void_p(self.s1.data_ptr() + 4 * off),
                    ctypes.c_void_p(self.s2.data_ptr() + 4 * off) if self.s2 is not None else None,
                    ctypes.byref(o), ctypes.c_void_p(stream)), "b2s_step_fused")
            if self.kind == 1 and o.first_step:
                for k in ks:
                    st[self.flat.params[k]]["momentum_buffer"] = self._view(self.s1, k)


def fused_step(self, mu=None, optimizer=None):
    """The rest of iter()'s minibatch body after ``comp_g`` (opt.py:622-659, 696-699) as device work without a host
    sync: penalty gradient when ``g > 0`` (opt.py:631-634), its norm clip (opt.py:539-542; the scale stays on the
    device), ``p = grad f + mu * sign * grad rho``, ``param.grad`` = views of its flat fp32 rounding, and the SGD /
    Adam update applied in place to the flat parameter vector (``b2s_clip_norm`` + ``b2s_step_fused``).  Optimizers
    other than plain ``torch.optim.SGD`` / ``Adam`` get the fused assembly followed by their own ``step()``."""
    import ctypes
    from . import _lib
    from .hvp_operator import flat_params_of
    lib = _lib.load()
    optimizer = optimizer if optimizer is not None else self.optimizer
    if mu is None:
        mu = self.mu(getattr(self, "i", 0)) if callable(self.mu) else self.mu
    flat = flat_params_of(self.model, self.device)
    fo = getattr(self, "_fused_opt", None)
    if fo is None or fo.optimizer is not optimizer or fo.flat is not flat:
        fo = self._fused_opt = _FusedOptimizer(optimizer, flat) if _FusedOptimizer.supported(optimizer, flat) else False
        self._step_p32 = torch.empty(flat.n, dtype=torch.float32, device=self.device)
        self._step_grad_views = [self._step_p32[o:o + n].view(q.shape) for q, o, n in zip(flat.params, flat.offsets, flat.sizes)]
        self._clip_scratch = torch.zeros(int(lib.b2s_clip_scratch_doubles()) + 2, dtype=torch.float64, device=self.device)
    if self.hvp_op.stored_grad is not None:                                   # opt.py:624-627
        self.gradf = self.hvp_op.stored_grad.data
    else:
        self.gradf = torch.zeros(self.ndim, dtype=torch.float64, device=self.device)
    st = torch.cuda.current_stream(self.device).cuda_stream
    coef, gr, scale2 = 0.0, None, None
    if self.g > 0:                                                            # opt.py:631-634
        self.gradrho = gr = self.hvp_op.vGHv(self.v, storedGrad=True)
        sign = 1 if self.rho > self.K else -1
        coef = float(mu) * sign
        if self.gradg_clip is not None:                                       # opt.py:539-542, scale kept on the device
            scale2 = self._clip_scratch[-2:]
            _lib.check(lib.b2s_clip_norm(ctypes.c_void_p(gr.data_ptr()), flat.n, ctypes.c_double(float(self.gradg_clip)),
                                         ctypes.c_void_p(self._clip_scratch.data_ptr()), ctypes.c_void_p(scale2.data_ptr()),
                                         ctypes.c_void_p(st)), "b2s_clip_norm")
    views = self._step_grad_views
    if flat.params[0].grad is not views[0]:
        for q, view in zip(flat.params, views):                               # opt.py:654-659 without slicing / casting kernels
            q.grad = view
    if fo:
        flat.attach()
        fo.step(lib, self.gradf, gr, coef, scale2, self._step_p32, st)
        if scale2 is not None:
            gr.mul_(scale2[1])                                                # self.gradrho *= clip / grn, on the device
    else:
        o = _lib.StepOpt()
        o.kind, o.write_gradrho = 0, 1
        _lib.check(lib.b2s_step_fused(ctypes.c_void_p(self.gradf.data_ptr()), ctypes.c_void_p(gr.data_ptr()) if gr is not None else None,
                                      ctypes.c_double(coef), ctypes.c_void_p(scale2.data_ptr()) if scale2 is not None else None,
                                      flat.n, None, ctypes.c_void_p(self._step_p32.data_ptr()), None, None, None,
                                      ctypes.byref(o), ctypes.c_void_p(st)), "b2s_step_fused")
        optimizer.step()
    if gr is not None:
        self.gradg = gr if coef >= 0 else -gr
    else:
        if getattr(self, "_zero64", None) is None or self._zero64.numel() != self.ndim:
            self._zero64 = torch.zeros(self.ndim, dtype=torch.float64, device=self.device)
        self.gradg = self._zero64                                             # opt.py:636


def _append_log(path, text, mode="a"):
    with open(path, mode) as fh:
        fh.write(text + "\n")


def iter_epoch(self):
    """``OptWBoundEignVal.iter`` (opt.py:580-763) for the power-iteration branch with a plain torch optimizer: same
    side effects (``f, g, h, rho, norm, gradf, gradg``, verbose log lines, scheduler step), with the minibatch body
    = ``comp_g`` + ``fused_step`` and the epoch loss through the forward-only evaluation pass.  Everything else
    (``pow_iter=False``, the K-FAC / Entropy-SGD / SAM optimizers with their extra forward passes) is handed to the
    reference's own ``iter``."""
    import random
    name = self.optimizer.__class__.__name__
    original = getattr(self, "_b200_reference_iter", None)
    if (not self.pow_iter or name in ("KFACOptimizer", "EntropySGD", "SAM")) and original is not None:
        return original()
    t_iter = time.time()
    self.model.train()
    if self.verbose:
        _append_log(self.verbose_log_file, "batch\t rho\t norm\t gradf\t gradg", "w" if self.i == 0 else "a")
    mu = self.mu(self.i) if callable(self.mu) else self.mu
    pick = random.randint(0, len(self.dataloader) - 1)            # batch of the end-of-epoch rho estimate (opt.py:604)
    t_g = t_gg = a0 = a1 = a2 = 0.0
    rdata = None
    for j, data in enumerate(self.dataloader):
        if j == pick:
            rdata = data
        t = time.time()
        self.comp_g(data)
        t_g += time.time() - t
        t = time.time()
        try:
            fused_step(self, mu)
        except RuntimeError:                                       # opt.py:696-699
            self.model_load("./models/" + self.header2 + "_trained_model.pt")
        t_gg += time.time() - t
        a0 += self.hvp_op.aTime0
        a1 += self.hvp_op.aTime1
        a2 += self.hvp_op.aTime2
        if self.verbose:
            _append_log(self.verbose_log_file, "%d\t %f\t %f\t %f\t %f" % (
                j, self.rho, self.norm, torch.norm(self.gradf.detach()), torch.norm(self.gradg.detach())))
        self.mem_check()
    t = time.time()
    fs, sizes = [], []
    for data in self.dataloader:                                   # opt.py:728-739
        inputs, target = self.prep_data(data)
        sizes.append(len(target))
        fs.append(self.comp_f(inputs, target)[0])
    self.f = np.average(fs, weights=sizes)
    self.comp_g(rdata)
    self.h = self.f + mu * self.g
    t_test = time.time() - t
    if self.verbose:
        old_stdout = sys.stdout
        with open(self.verbose_log_file, "a") as log_file:
            sys.stdout = log_file
            try:
                for head, val in (("G ", t_g), ("Grad G ", t_gg), ("Test ", t_test), ("Iteration ", time.time() - t_iter),
                                  ("Autograd 0 ", a0), ("Autograd 1 ", a1), ("Autograd 2 ", a2), ("K-FAC ", self.kTime)):
                    _time_hms(val, head)
            finally:
                sys.stdout = old_stdout
    if self.scheduler is not None and self.scheduler.__class__.__name__ == "ReduceLROnPlateau":   # opt.py:760-763
        self.scheduler.step(self.f)
    elif self.scheduler is not None:
        self.scheduler.step()


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
        self.gradg = torch.zeros(self.ndim, dtype=torch.float64, device=self.device)
        self.kTime = 0
        self.optimizer, self.scheduler, self.dataloader = None, None, None
        self.pow_iter, self.i, self.f, self.h, self.header2 = True, 0, 0, 0, "spectral"

    def random_v(self):   # opt.py:324-325
        return torch.from_numpy(1.0 / np.sqrt(self.ndim) * np.ones(self.ndim)).to(self.device)

    comp_rho = comp_rho
    comp_gradrho = comp_gradrho
    comp_g = comp_g
    comp_f = comp_f
    iter = iter_epoch

    def prep_data(self, data):    # opt.py:338-346
        if isinstance(data, (list, tuple)):
            inputs, target = data
        elif isinstance(data, dict):
            inputs, target = data["image"], data["label"]
        else:
            raise Exception("Data type not supported")
        return inputs.to(self.device), target.to(self.device)

    def mem_check(self):
        pass

    def model_load(self, path):
        raise RuntimeError("SpectralState has no checkpoint to fall back to (%s)" % path)

    def kfac(self, r):            # opt.py:384-416
        from .kfac import kfac as _kfac
        return _kfac(self, r)

    def init_kfac(self, data):    # opt.py:362-382
        from .kfac import init_kfac as _init
        return _init(self, data)

    assemble_step = assemble_step
    rho_test = rho_test

    fused_step = fused_step

    def regularized_step(self, data, optimizer):
        """One minibatch of iter() (opt.py:608-699, the pow_iter branch with a plain optimizer): comp_g, grad f,
        penalty gradient when g > 0 with its clip, step assembly into param.grad and the optimizer update, the last
        three fused on the device (``fused_step``)."""
        self.comp_g(data)
        self.fused_step(optimizer=optimizer)

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
