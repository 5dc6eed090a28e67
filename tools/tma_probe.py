"""Probe of the TMA-fed tcgen05 contraction (run under gpurun): a 1x1 convolution followed by a 3x3
convolution on small images; the forward output of the second convolution and the adjoint of its input
(its dgrad) from the CUDA library are compared with an fp64 torch evaluation of the same fp32 inputs.
Usage: tma_probe.py [W] [Cin] [Cout] [batch].  Test infrastructure."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from optwboundeigenval_b200 import _lib                                    # noqa: E402
from optwboundeigenval_b200.hvp_operator import B200HVPOperator, clear_plans   # noqa: E402


class TwoConv(nn.Module):
    def __init__(self, cin, cout, k, stride=1, pad=None):
        super().__init__()
        self.c1 = nn.Conv2d(16, cin, 1, bias=False)
        self.c2 = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2 if pad is None else pad, bias=False)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cout, 10)

    def forward(self, x):
        return self.fc(torch.flatten(self.pool(self.c2(self.c1(x))), 1))


def read_tensor(plan, adjoint, order, t, batch):
    vt = plan.tape.tensors[t]
    out = np.zeros((batch,) + tuple(vt.shape), dtype=np.float32)
    _lib.check(plan.lib.b2s_debug_read(plan.handle, adjoint, order, t, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def probe(W=32, cin=48, cout=12, batch=4, k=3, verbose=True, stride=1, pad=None):
    """returns {(mode, what): relative L2 error against fp64} for what in fwd / dgrad / wgrad, mode in 0 (CUDA cores) / 2 (tensor cores)"""
    lib = _lib.load()
    errs = {}
    torch.manual_seed(3)
    m = TwoConv(cin, cout, k, stride, pad).train()
    x = torch.randn(batch, 16, W, W)
    y = torch.randint(0, 10, (batch,))
    md = TwoConv(cin, cout, k, stride, pad).double()
    md.load_state_dict({k_: v.double() for k_, v in m.state_dict().items()})
    h1 = md.c1(x.double())
    h1.retain_grad()
    h2 = md.c2(h1)
    out = md.fc(torch.flatten(md.pool(h2), 1))
    nn.CrossEntropyLoss()(out, y).backward()
    want_f = h2.detach().numpy()
    want_b = h1.grad.numpy()
    want_w = md.c2.weight.grad.numpy()
    w_off = md.c1.weight.numel()
    for mode in (0, 2):
        clear_plans()
        _lib.check(lib.b2s_set_tensor_core_mode(mode))
        op = B200HVPOperator(m, [x, y], nn.CrossEntropyLoss())
        grad = op.prepare_grad()
        torch.cuda.synchronize()
        plan = op.plan
        convs = [o for o in plan.tape.ops if o.kind == 1]
        got_f = read_tensor(plan, 0, 0, convs[1].out, batch).astype(np.float64)
        got_b = read_tensor(plan, 1, 0, convs[0].out, batch).astype(np.float64)
        got_w = grad.cpu().numpy()[w_off:w_off + want_w.size].reshape(want_w.shape)
        for name, got, want in (("fwd", got_f, want_f), ("dgrad", got_b, want_b), ("wgrad", got_w, want_w)):
            d = got - want
            errs[(mode, name)] = float(np.linalg.norm(d) / np.linalg.norm(want))
            if verbose:
              print("W=%d Cin=%d Cout=%d k=%d batch=%d mode=%d %-5s rel_l2 %.3e  max_rel %.3e" % (
                W, cin, cout, k, batch, mode, name, np.linalg.norm(d) / np.linalg.norm(want),
                float(np.abs(d).max() / np.abs(want).mean())), flush=True)
            if verbose and name != "wgrad" and mode == 2 and np.linalg.norm(d) / np.linalg.norm(want) > 1e-4:
                # where is it wrong?  per output channel / per pixel-row error
                e = np.abs(d)
                print("   err by channel:", np.round(e.mean(axis=(0, 2, 3)) / np.abs(want).mean(), 4)[:16])
                print("   err by row    :", np.round(e.mean(axis=(0, 1, 3)) / np.abs(want).mean(), 4)[:32])
                print("   err by col    :", np.round(e.mean(axis=(0, 1, 2)) / np.abs(want).mean(), 4)[:32])
                print("   err by image  :", np.round(e.mean(axis=(1, 2, 3)) / np.abs(want).mean(), 4)[:8])
    _lib.check(lib.b2s_set_tensor_core_mode(1))
    clear_plans()
    return errs


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    probe(*a)
