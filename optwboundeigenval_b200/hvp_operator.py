"""B200HVPOperator -- drop-in for ``opt.HVPOperator`` (reference opt.py:48-192).

Same constructor, attributes and methods; the arithmetic runs in
libb200spectral (hand-written sm_100a CUDA) instead of nested
``torch.autograd.grad`` calls.  PyTorch is used for device memory, streams and
parameter storage only.

    opt.HVPOperator(model, data, criterion, use_gpu=True, mem_track=False)
        .Hv(vec, storedGrad=False)      -> fp64 tensor [P]        opt.py:77-108
        .vGHv(vec, storedGrad=False)    -> fp64 tensor [P]        opt.py:110-152
        .prepare_grad()                 -> fp64 tensor [P]        opt.py:175-192
        .zero_grad(model=None), .prep_data(data), .mem_check()
        .stored_grad .size .aTime0 .aTime1 .aTime2 .device .mem_max

Differences that are deliberate:
  * no CPU path: without CUDA (or without the built library) construction raises;
  * ``stored_grad`` stays on the device (the reference parks it on the CPU and
    copies it back for every Hv, opt.py:87,91); callers only ever do
    ``.data.to(device)`` on it (opt.py:624-625), which works unchanged;
  * ``vGHv`` may be called more than once (the reference frees its graph).
"""
from __future__ import annotations

import ctypes
import os
import time
import weakref
from typing import Optional

import numpy as np
import torch

from . import _lib
from .tracer import HEAD_CE, HEAD_SIGMOID_WBCE, HEAD_SOFTMAX_CE, HEAD_WBCE, Tape, trace


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("optwboundeigenval_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class PowerOutcome:
    __slots__ = ("v", "lam", "norm", "rn", "vnn", "iters", "converged", "stop", "trajectory")


class SpectralPlan:
    """Python owner of one ``b2s_plan`` (tape + workspaces) for a (model, loss, input shape)."""

    def __init__(self, model, criterion, input_shape, max_batch: int, device: torch.device, use_graphs: bool = True):
        _require_cuda()
        self.lib = _lib.load()
        self.tape: Tape = trace(model, criterion, input_shape)
        self.device = device
        self.max_batch = int(max_batch)
        self.P = self.tape.n_params
        handle = ctypes.c_void_p()
        tens = self.tape.c_tensors()
        ops = self.tape.c_ops()
        bufs = (ctypes.c_int64 * len(self.tape.buf_elems))(*self.tape.buf_elems)
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(self.lib.b2s_plan_create(tens, len(tens), bufs, len(bufs), ops, len(ops), self.tape.logits,
                                            self.tape.head, self.P, self.max_batch, dev_index,
                                            ctypes.byref(handle)), "b2s_plan_create")
        self.handle = handle
        self.dev_index = dev_index
        if os.environ.get("B2S_NO_GRAPHS"):          # debugging aid: eager launches (kernel traces, sanitizers)
            use_graphs = False
        _lib.check(self.lib.b2s_plan_set_graphs(self.handle, 1 if use_graphs else 0))
        self._bn_ptrs = None
        self.world = 1
        self.rank = 0
        self.owner = None                 # weakref of the operator whose base pass is cached (see B200HVPOperator._own)
        self.global_batch = 0
        self._vin = None                  # staging vector for host inputs (fp64, device)
        self._finalizer = weakref.finalize(self, SpectralPlan._destroy, self.lib, handle)

    @staticmethod
    def _destroy(lib, handle):
        try:
            lib.b2s_plan_destroy(handle)
        except Exception:   # noqa: BLE001  (interpreter shutdown)
            pass

    # ---- helpers ----------------------------------------------------------------------------
    def _bind_stream(self):
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.b2s_plan_set_stream(self.handle, ctypes.c_void_p(st)))

    def _bind_bn(self):
        ptrs = tuple((m.running_mean.data_ptr(), m.running_var.data_ptr()) for m in self.tape.bn_modules)
        if ptrs != self._bn_ptrs:
            for slot, (rm, rv) in enumerate(ptrs):
                _lib.check(self.lib.b2s_plan_set_bn_buffers(self.handle, slot, ctypes.c_void_p(rm), ctypes.c_void_p(rv)))
            self._bn_ptrs = ptrs

    def workspace_bytes(self) -> int:
        return int(self.lib.b2s_plan_workspace_bytes(self.handle))

    @staticmethod
    def wbce_coefficients(target: torch.Tensor, sums=None):
        """Per-entry weights of W_BCEWithLogitsLoss divided by the per-class valid counts and the number
        of classes that have a valid label (dcnn.py:386-400), with NaN labels masked.  ``sums`` lets the
        data-parallel path substitute batch-global counts."""
        valid = ~torch.isnan(target)
        t = torch.where(valid, target, torch.zeros_like(target))
        p = torch.trunc(t.double().sum())
        s = valid.double().sum()
        n_c = valid.double().sum(0)
        if sums is not None:
            p, s, n_c = sums(p, s, n_c)
        degenerate = (p == 0) | (p == s)
        safe_p = torch.where(p == 0, torch.ones_like(p), p)
        safe_sp = torch.where(s - p == 0, torch.ones_like(p), s - p)
        pos_w = (s / safe_p)
        neg_w = (s / safe_sp)
        w = torch.where(degenerate, t + 1.0, t * (pos_w - neg_w).float() + neg_w.float())
        class_ok = n_c > 0
        c_valid = class_ok.double().sum().clamp_min(1.0)
        denom = (n_c.clamp_min(1.0) * c_valid).float()
        coef = torch.where(valid, w / denom, torch.zeros_like(w))
        return t.contiguous(), coef.contiguous()

    # ---- passes -----------------------------------------------------------------------------
    def _head_inputs(self, target, batch, gbatch, use_global):
        head = self.tape.head
        if head in (HEAD_CE, HEAD_SOFTMAX_CE):
            nc = self.tape.tensors[self.tape.logits].numel
            return target.to(self.device, torch.int64).contiguous().view(-1), None, self.tape.head_scale(gbatch, nc)
        t = target.to(self.device, torch.float32)
        if t.dim() == 1:
            t = t.view(-1, 1)
        y, coef = self.wbce_coefficients(t, self._global_sums if use_global else None)
        return y, coef, 1.0

    def eval_pass(self, params: torch.Tensor, x: torch.Tensor, target: torch.Tensor):
        """comp_f (opt.py:544-572): forward only, evaluation-mode BatchNorm; returns (loss fp64 [1], model output
        [batch, classes] fp32) device tensors -- the output with the model's own softmax / sigmoid tail applied,
        as ``self.model(inputs)`` returns it.  Local to this rank; the cached base pass is forgotten."""
        batch = int(x.shape[0])
        if batch > self.max_batch:
            raise RuntimeError("batch %d exceeds the plan's max_batch %d" % (batch, self.max_batch))
        self._bind_stream()
        self._bind_bn()
        x = x.to(self.device, torch.float32).contiguous()
        y, coef, scale = self._head_inputs(target, batch, batch, False)
        nc = self.tape.tensors[self.tape.logits].numel
        logits = torch.empty(batch, nc, dtype=torch.float32, device=self.device)
        loss = torch.empty(1, dtype=torch.float64, device=self.device)
        self.owner = None
        _lib.check(self.lib.b2s_eval_pass(self.handle, _ptr(params), _ptr(x), _ptr(y), _ptr(coef), batch, scale,
                                          _ptr(logits), _ptr(loss)), "b2s_eval_pass")
        self._keep = (params, x, y, coef)
        if self.tape.kl_reduction is not None:
            logits = torch.log_softmax(logits, dim=1)
        elif self.tape.head == HEAD_SOFTMAX_CE:
            logits = torch.softmax(logits, dim=1)
        elif self.tape.head == HEAD_SIGMOID_WBCE:
            logits = torch.sigmoid(logits)
        return loss, logits

    def base_pass(self, params: torch.Tensor, x: torch.Tensor, target: torch.Tensor):
        """forward + loss + gradient; returns (grad fp64 [P], loss fp64 [1]) device tensors."""
        batch = int(x.shape[0])
        if batch > self.max_batch:
            raise RuntimeError("batch %d exceeds the plan's max_batch %d" % (batch, self.max_batch))
        self._bind_stream()
        self._bind_bn()
        x = x.to(self.device, torch.float32).contiguous()
        gbatch = batch
        if self.world > 1:
            # shards may be ragged (the last minibatch of a per-rank loader): weigh by the global sample count
            import torch.distributed as dist
            nb = torch.tensor([batch], dtype=torch.int64, device=self.device)
            dist.all_reduce(nb)
            gbatch = int(nb.item())
            _lib.check(self.lib.b2s_plan_set_global_batch(self.handle, gbatch), "b2s_plan_set_global_batch")
        self.global_batch = gbatch
        y, coef, scale = self._head_inputs(target, batch, gbatch, self.world > 1)
        grad = torch.empty(self.P, dtype=torch.float64, device=self.device)
        loss = torch.empty(1, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.b2s_base_pass(self.handle, _ptr(params), _ptr(x), _ptr(y), _ptr(coef), batch, scale,
                                          _ptr(grad), _ptr(loss)), "b2s_base_pass")
        self._keep = (params, x, y, coef)      # inputs must outlive the asynchronous copies
        return grad, loss

    def _global_sums(self, p, s, n_c):
        import torch.distributed as dist
        buf = torch.cat([p.view(1), s.view(1), n_c.view(-1)])
        dist.all_reduce(buf)
        return buf[0], buf[1], buf[2:]

    def hv(self, v: torch.Tensor) -> torch.Tensor:
        self._bind_stream()
        out = torch.empty(self.P, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.b2s_hv(self.handle, _ptr(v), _ptr(out)), "b2s_hv")
        return out

    def vghv(self, v: torch.Tensor) -> torch.Tensor:
        self._bind_stream()
        out = torch.empty(self.P, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.b2s_vghv(self.handle, _ptr(v), _ptr(out)), "b2s_vghv")
        return out

    def power_iterate(self, v0: torch.Tensor, eps: float, max_iter: int, alphas=None, want_trajectory=False,
                      precond=False):
        self._bind_stream()
        v = v0.to(self.device, torch.float64).contiguous().clone()
        cfg = _lib.PowerCfg()
        cfg.max_iter = int(max_iter)
        cfg.eps = float(eps)
        cfg.precond = 1 if precond else 0
        keep = None
        if alphas is not None:
            keep = (ctypes.c_double * int(max_iter))(*[float(a) for a in alphas])
            cfg.h_alpha = ctypes.cast(keep, ctypes.POINTER(ctypes.c_double))
        res = _lib.PowerResult()
        traj = np.zeros((max(int(max_iter), 1), 5), dtype=np.float64) if want_trajectory else None
        _lib.check(self.lib.b2s_power_iterate(self.handle, _ptr(v), ctypes.byref(cfg), ctypes.byref(res),
                                              traj.ctypes.data_as(ctypes.c_void_p) if traj is not None else None),
                   "b2s_power_iterate")
        out = PowerOutcome()
        out.v = v
        out.lam, out.norm, out.rn, out.vnn = res.lam, res.norm, res.rn, res.vnn
        out.iters, out.converged = int(res.iters), bool(res.converged)
        out.stop = [res.stop[0], res.stop[1], res.stop[2]]
        out.trajectory = traj[:out.iters + 1] if traj is not None else None
        return out

    def set_bn_third_order(self, exact: bool):
        """False (default): reproduce nested torch.autograd's third-order BatchNorm result; True: exact."""
        _lib.check(self.lib.b2s_plan_set_bn_third_order(self.handle, 1 if exact else 0))

    def profile(self, order: int, reps: int = 3, raw: bool = False):
        """Per kernel family: launches / ms / algorithmic FLOPs and bytes of one pass (CUDA events on the
        launching stream around every kernel)."""
        self._bind_stream()
        cap = 4096 if raw else 64
        arr = (_lib.ProfEntry * cap)()
        n = ctypes.c_int32(0)
        _lib.check(self.lib.b2s_profile_pass(self.handle, order | (0x100 if raw else 0), reps, arr, cap,
                                             ctypes.byref(n)), "b2s_profile_pass")
        return [dict(name=arr[i].name.decode(), launches=arr[i].launches, ms=arr[i].ms, flops=arr[i].flops,
                     bytes=arr[i].bytes) for i in range(n.value)]

    # ---- data parallelism ---------------------------------------------------------------------
    def init_comm(self):
        """One NCCL communicator per plan, bootstrapped through torch.distributed (already initialised)."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_char * 128)()
            _lib.check(self.lib.b2s_comm_unique_id(buf), "b2s_comm_unique_id")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(self.device)
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().numpy().tobytes())
        _lib.check(self.lib.b2s_comm_init(self.handle, ctypes.c_char_p(raw), rank, world), "b2s_comm_init")
        self.world, self.rank = world, rank
        # exchange buffers for the one-shot peer all-reduce of the per-layer BatchNorm sums (NVLink peer memory)
        if dist.get_backend() == "nccl" and world <= 8:
            mine = (ctypes.c_char * 64)()
            _lib.check(self.lib.b2s_comm_peer_local(self.handle, mine), "b2s_comm_peer_local")
            h = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).clone().to(self.device)
            hs = [torch.empty_like(h) for _ in range(world)]
            dist.all_gather(hs, h)
            allh = b"".join(bytes(t.cpu().numpy().tobytes()) for t in hs)
            _lib.check(self.lib.b2s_comm_peer_attach(self.handle, ctypes.c_char_p(allh)), "b2s_comm_peer_attach")
            # every rank or none: a rank that could not map its peers would use NCCL while the others poll flags
            ok = torch.tensor([int(self.lib.b2s_comm_peer_ready(self.handle))], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                _lib.check(self.lib.b2s_comm_peer_disable(self.handle), "b2s_comm_peer_disable")


# one plan per live model (keyed by identity); new operators per minibatch reuse it (opt.py:424
# constructs a new HVPOperator for every batch)
_PLANS = {}
_DATA_PARALLEL = True


def set_data_parallel(on: bool):
    """True (default): a minibatch is sharded over the ranks of torch.distributed (synced BatchNorm sums, all-reduced
    results).  False: *replicas only* -- every rank works on its own minibatches with no collective at all (the
    ``rho_test`` sweep, opt.py:882-910).  Changing the mode drops the cached plans."""
    global _DATA_PARALLEL
    if bool(on) != _DATA_PARALLEL:
        _DATA_PARALLEL = bool(on)
        _PLANS.clear()


def plan_for(model, criterion, x: torch.Tensor, device, max_batch=None) -> SpectralPlan:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:        # the reference says torch.device('cuda') (opt.py:247)
        device = torch.device("cuda", torch.cuda.current_device())
    shape = tuple(x.shape[1:])
    key = (id(model), criterion.__class__.__name__, getattr(criterion, "reduction", None), shape)
    entry = _PLANS.get(key)
    batch = int(x.shape[0])
    if entry is not None:
        ref, plan = entry
        if ref() is model and plan.max_batch >= batch and plan.device == device:
            return plan
    plan = SpectralPlan(model, criterion, shape, max(batch, max_batch or 0), device)
    if _DATA_PARALLEL and torch.distributed.is_available() and torch.distributed.is_initialized():
        plan.init_comm()
    _PLANS[key] = (weakref.ref(model), plan)
    return plan


def clear_plans():
    _PLANS.clear()


class FlatParams(object):
    """One persistent flat fp32 vector per model in ``model.parameters()`` order (the layout of every spectral
    vector, opt.py:102,191).

    ``gather()`` refreshes it with ONE multi-tensor copy instead of a ``torch.cat`` of 119..364 tensors per
    minibatch.  ``attach()`` goes further and makes every ``param.data`` a view INTO the vector (what
    DistributedDataParallel does with its buckets): gathering becomes free and the fused optimizer step
    (``spectral.fused_step``) updates the parameters in place on the flat vector.  Parameter objects, their
    identity and ``state_dict`` keys are unchanged; an attachment broken from outside (``param.data = ...``,
    ``model.to(...)``) is detected by pointer comparison and redone."""

    def __init__(self, model, device):
        self.params = [q for q in model.parameters()]
        self.sizes = [q.numel() for q in self.params]
        self.n = sum(self.sizes)
        self.flat = torch.empty(self.n, dtype=torch.float32, device=device)
        self.views, self.offsets = [], []
        i = 0
        for q, n in zip(self.params, self.sizes):
            self.views.append(self.flat[i:i + n].view(q.shape))
            self.offsets.append(i)
            i += n
        self.attached = False

    def matches(self, model):
        ps = list(model.parameters())
        return len(ps) == len(self.params) and all(a is b for a, b in zip(ps, self.params))

    def is_attached(self):
        base, es = self.flat.data_ptr(), self.flat.element_size()
        return self.attached and all(q.dtype == torch.float32 and q.data_ptr() == base + o * es and q.is_contiguous()
                                     for q, o in zip(self.params, self.offsets))

    def gather(self) -> torch.Tensor:
        if not self.is_attached():
            self.attached = False
            torch._foreach_copy_(self.views, [q.detach() for q in self.params])
        return self.flat

    def attach(self) -> torch.Tensor:
        if not self.is_attached():
            with torch.no_grad():
                torch._foreach_copy_(self.views, [q.detach() for q in self.params])
                for q, v in zip(self.params, self.views):
                    q.data = v
            self.attached = True
        return self.flat


_FLAT = weakref.WeakKeyDictionary()


def flat_params_of(model, device) -> FlatParams:
    fp = _FLAT.get(model)
    if fp is None or fp.flat.device != device or not fp.matches(model):
        fp = _FLAT[model] = FlatParams(model, device)
    return fp


def flat_parameters(model) -> torch.Tensor:
    dev = next(model.parameters()).device
    return flat_params_of(model, dev).gather()


class B200HVPOperator(object):
    """See module docstring; mirrors opt.py:48-192."""

    def __init__(self, model, data, criterion, use_gpu=True, mem_track=False):
        _require_cuda()
        _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        # opt.py:56 moves the model; when every parameter and buffer already lives on the device that is a no-op which
        # still walks all modules (1.8 ms per minibatch for DenseNet3, on the critical path of the regularised step)
        if any(t.device != self.device for t in model.parameters()) or any(t.device != self.device for t in model.buffers()):
            model = model.to(self.device)
        self.model = model
        self.data = data
        self.criterion = criterion
        self.use_gpu = True
        self.stored_grad = None
        self.mem_track = mem_track
        self.mem_max = 0
        self.size = 0
        self.aTime0 = self.aTime1 = self.aTime2 = 0
        self.plan: Optional[SpectralPlan] = None
        self.loss_value = None
        # True: a PINNED host vector is copied to the device without waiting on the host (the caller must not
        # overwrite it before the call's result has been synchronised); False: torch's blocking semantics
        self.async_host_vectors = False

    # -- housekeeping, same behaviour as the reference (opt.py:72-75,154-173) --
    def mem_check(self):
        if self.mem_track:
            self.mem_max = np.max([self.mem_max, torch.cuda.memory_allocated()])

    def zero_grad(self, model=None):
        if model is None:
            model = self.model
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        if grads:                                     # one multi-tensor launch instead of one kernel per parameter
            torch._foreach_zero_(grads)

    def prep_data(self, data):
        if type(data) == list or type(data) == tuple:
            inputs, target = data
        elif type(data) == dict:
            inputs, target = data["image"], data["label"]
        else:
            raise Exception("Data type not supported")
        return inputs.to(self.device), target.to(self.device)

    def _vec(self, vec):
        if type(vec) is np.ndarray:
            vec = torch.from_numpy(vec)
        if vec.is_cuda and vec.dtype == torch.float64 and vec.is_contiguous():
            return vec
        if not vec.is_cuda and vec.dtype == torch.float64 and self.plan is not None and vec.numel() == self.plan.P:
            # host vector: one copy into the plan's staging vector (no allocation, no second cast kernel)
            if self.plan._vin is None:
                self.plan._vin = torch.empty(self.plan.P, dtype=torch.float64, device=self.device)
            self.plan._vin.copy_(vec.reshape(-1), non_blocking=self.async_host_vectors and vec.is_pinned())
            return self.plan._vin
        return vec.to(self.device).double().contiguous()

    # -- the three entry points --
    def prepare_grad(self):
        inputs, target = self.prep_data(self.data)
        self.size = len(target)
        if not self.model.training:
            raise RuntimeError("B200HVPOperator: the model must be in train() mode (comp_rho forces it, opt.py:421)")
        self.plan = plan_for(self.model, self.criterion, inputs, self.device)
        start = time.time()
        grad, loss = self.plan.base_pass(flat_parameters(self.model), inputs, target)
        nbt = [m.num_batches_tracked for m in self.plan.tape.bn_modules if m.num_batches_tracked is not None]
        if nbt:                                       # train-mode forward side effect
            torch._foreach_add_(nbt, 1)
        self.plan.owner = weakref.ref(self)
        self.aTime0 += time.time() - start
        self.loss_value = loss
        return grad

    def _ensure_grad(self, storedGrad):
        if not (storedGrad and self.stored_grad is not None):
            self.zero_grad()
            self.stored_grad = self.prepare_grad()
        elif self.plan.owner is None or self.plan.owner() is not self:
            self._own()

    def _own(self):
        """All operators of one (model, loss, input shape) share a plan, whose caches hold the LAST base pass.  The
        reference's operators are independent objects (each keeps its own graph, opt.py:175-192), so when another
        operator -- or init_kfac -- ran a base pass in between, this operator's minibatch is passed again before its
        Hv / vGHv; the train-mode side effects of that repeat (BatchNorm running statistics, num_batches_tracked)
        are undone, because the reference would not have run a second forward."""
        bns = self.plan.tape.bn_modules
        saved = [(m.running_mean.clone(), m.running_var.clone(),
                  m.num_batches_tracked.clone() if m.num_batches_tracked is not None else None) for m in bns]
        a0 = self.aTime0
        self.prepare_grad()
        self.aTime0 = a0
        with torch.no_grad():
            for m, (rm, rv, nb) in zip(bns, saved):
                m.running_mean.copy_(rm)
                m.running_var.copy_(rv)
                if nb is not None:
                    m.num_batches_tracked.copy_(nb)

    def Hv(self, vec, storedGrad=False):
        self._ensure_grad(storedGrad)
        vec = self._vec(vec)
        self.mem_check()
        start = time.time()
        out = self.plan.hv(vec)
        self.aTime1 += time.time() - start
        self.mem_check()
        return out

    def vGHv(self, vec, storedGrad=False):
        self._ensure_grad(storedGrad)
        vec = self._vec(vec)
        self.mem_check()
        start = time.time()
        out = self.plan.vghv(vec)
        self.aTime2 += time.time() - start
        self.mem_check()
        return out

    # -- fused loop used by the comp_rho replacement (spectral.py) --
    def power_iterate(self, v0, eps, max_iter, alphas=None, want_trajectory=False):
        self._ensure_grad(True)
        start = time.time()
        out = self.plan.power_iterate(self._vec(v0), eps, max_iter, alphas, want_trajectory)
        self.aTime1 += time.time() - start
        return out
