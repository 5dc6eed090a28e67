"""Parity modulo fp32-ambiguous ReLU decisions (test infrastructure).

The gradient and the Hessian-vector product of a ReLU network are discontinuous where a pre-activation
crosses zero.  A minibatch of DenseNet3 / VGG16 evaluates 10^6..10^7 ReLU inputs, so a handful of them
always lie within fp32 rounding of zero; two correct fp32 implementations with a different summation order
(the reference's oneDNN/cuDNN kernels, this library's CUDA-core kernels, its tcgen05 kernels) may decide
such an element differently, and ONE flipped decision moves the gradient by more than the rtol 1e-4 the
north star asks for (the reference's own fp32 result is that far from an fp64 evaluation of itself,
tests/test_gpu_parity.py::test_full_size_chest_models_against_cpu_autograd).

Max pooling is piecewise linear too: a window whose two largest elements differ by less than fp32 rounding may route
its adjoint to either element.

`explain_by_kinks` makes that statement checkable instead of loosening the tolerance:
  1. the fp64 jet oracle (oracle/jet_oracle.py, pinned to nested autograd by tests/test_jet_oracle.py)
     evaluates the same tape; every ReLU decision the GPU took differently must sit on a pre-activation
     whose fp64 magnitude is below `kink_tol` x the rms of that tensor (i.e. undecidable in fp32);
  2. the oracle re-evaluated WITH the GPU's decisions must reproduce the GPU's gradient / Hv / vGHv
     within the full rtol.
"""
import ctypes

import numpy as np
import torch

from conftest import rel_err
from optwboundeigenval_b200 import _lib, tracer
from optwboundeigenval_b200.hvp_operator import SpectralPlan
from oracle import autograd_oracle as ao
from oracle.jet_oracle import JetTapeOracle


def _read_value(plan, t, batch):
    vt = plan.tape.tensors[t]
    out = np.zeros((batch,) + tuple(vt.shape), dtype=np.float32)
    _lib.check(plan.lib.b2s_debug_read(plan.handle, 0, 0, t, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def gpu_relu_decisions(plan, batch):
    """op index -> bool tensor of the on/off decisions of the cached base pass (post-ReLU value > 0)"""
    out = {}
    for oi, op in enumerate(plan.tape.ops):
        if op.kind == tracer.OP_RELU or (op.flags & tracer.F_RELU):
            out[oi] = torch.from_numpy(_read_value(plan, op.out, batch) > 0)
    return out


def gpu_argmax_decisions(plan, batch):
    """op index -> the window element every max pool of the cached base pass selected, recomputed from the GPU's own
    fp32 input values with the same first-maximum rule (elementwise.cu maxpool_fwd_kernel / ATen)"""
    import torch.nn.functional as F
    out = {}
    for oi, op in enumerate(plan.tape.ops):
        if op.kind == tracer.OP_MAXPOOL:
            kh, kw, sh, sw, ph, pw = op.geom
            xin = torch.from_numpy(_read_value(plan, op.inp, batch))
            out[oi] = F.max_pool2d(xin, (kh, kw), (sh, sw), (ph, pw), return_indices=True)[1]
    return out


def _oracle(plan, model, x, y):
    tape = plan.tape
    x, y = x.detach().cpu(), y.detach().cpu()
    params = ao.flat_params(model).detach().cpu()       # the operator may have moved the model to the GPU
    if tape.head in (tracer.HEAD_WBCE, tracer.HEAD_SIGMOID_WBCE):
        t, coef = SpectralPlan.wbce_coefficients(y.float())
        return JetTapeOracle(tape, params, x, t, coef, loss_scale=1.0)
    return JetTapeOracle(tape, params, x, y)


def explain_by_kinks(op, model, x, y, checks, rtol=1e-4, kink_tol=1e-5, max_flip_fraction=1e-5):
    """`op`: B200HVPOperator whose base pass for (x, y) is cached.  `checks`: list of
    (name, kind, v, gpu_vector) with kind in {"grad", "hv", "vghv"} (v ignored for "grad").
    Returns the number of flipped decisions; raises AssertionError when the mismatch is not explained."""
    plan = op.plan
    batch = x.shape[0]
    decisions = gpu_relu_decisions(plan, batch)
    nat = _oracle(plan, model, x, y)
    nat.forward(0)
    flips = total = 0
    for oi, m_gpu in decisions.items():
        m_ref = nat.masks[oi]
        total += m_ref.numel()
        d = m_gpu != m_ref
        n = int(d.sum())
        if n:
            pre = nat.preact[oi]
            scale = float(pre.pow(2).mean().sqrt())
            worst = float(pre[d].abs().max())
            assert worst <= kink_tol * scale, (
                "op %d (%s): a ReLU decision differs where the pre-activation is NOT fp32-ambiguous: |pre| = %.3e, "
                "tensor rms %.3e" % (oi, plan.tape.ops[oi].name, worst, scale))
            flips += n
    assert flips <= max(1, int(max_flip_fraction * total)), "%d of %d ReLU decisions differ" % (flips, total)
    # max pooling is the other piecewise-linear op: a window whose two largest elements are within fp32 rounding of each
    # other may route its adjoint to either of them
    argmax = gpu_argmax_decisions(plan, batch)
    for oi, idx_gpu in argmax.items():
        idx_ref = nat.argmax[oi]
        d = idx_gpu != idx_ref
        n = int(d.sum())
        if n:
            xin = nat.view(nat.fw, 0, plan.tape.ops[oi].inp).flatten(2)
            a = xin.gather(2, idx_gpu.flatten(2)).view_as(d)[d]
            b = xin.gather(2, idx_ref.flatten(2)).view_as(d)[d]
            scale = float(xin.pow(2).mean().sqrt())
            worst = float((a - b).abs().max())
            assert worst <= kink_tol * scale, (
                "op %d (%s): a max-pool selection differs where the two candidates are NOT fp32-ambiguous: %.3e apart, "
                "tensor rms %.3e" % (oi, plan.tape.ops[oi].name, worst, scale))
            flips += int(((a != 0) | (b != 0)).sum())       # ties among zeros (behind a ReLU) carry no adjoint
    cond = _oracle(plan, model, x, y)
    cond.mask_override = decisions
    cond.argmax_override = argmax
    grad = cond.run(0)
    for name, kind, v, got in checks:
        if kind == "grad":
            want = grad
        elif kind == "hv":
            want = cond.run(1, torch.as_tensor(v, dtype=torch.float64)).clone()
        elif kind == "vghv":
            want = cond.vghv(torch.as_tensor(v, dtype=torch.float64))
        else:
            raise ValueError(kind)
        e = rel_err(np.asarray(got), want.numpy())
        assert e < rtol, "%s: %.3e from the oracle conditioned on the GPU's %d ambiguous ReLU decisions" % (name, e, flips)
    return flips
