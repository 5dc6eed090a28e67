"""Debug aid (run under gpurun): runs base pass + Hv of one config twice, tensor cores off and on, and
reports for every cached tensor (values and adjoints, orders 0 and 1) where the two runs differ."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from optwboundeigenval_b200 import _lib, zoo                                       # noqa: E402
from optwboundeigenval_b200.hvp_operator import B200HVPOperator, clear_plans       # noqa: E402


def read_tensor(plan, adjoint, order, t, batch):
    vt = plan.tape.tensors[t]
    out = np.zeros((batch,) + tuple(vt.shape), dtype=np.float32)
    _lib.check(plan.lib.b2s_debug_read(plan.handle, adjoint, order, t, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    lib = _lib.load()
    model, loss = zoo.build(kind)
    model.train()
    x, y = zoo.synthetic_batch(kind, batch)
    P = sum(p.numel() for p in model.parameters())
    v = torch.from_numpy(np.ones(P) / np.sqrt(P))
    snaps = {}
    for mode in (0, 1):
        clear_plans()
        _lib.check(lib.b2s_set_tensor_core_mode(mode))
        op = B200HVPOperator(model, [x, y], loss)
        op.Hv(v, storedGrad=True)
        torch.cuda.synchronize()
        plan = op.plan
        snap = {}
        for oi, o in enumerate(plan.tape.ops):
            for adj in (0, 1):
                for order in (0, 1):
                    snap[(oi, adj, order)] = read_tensor(plan, adj, order, o.out, batch)
        snaps[mode] = snap
        ops = plan.tape.ops
    names = {1: "conv", 2: "bn", 3: "relu", 4: "maxpool", 5: "avgpool", 6: "copy"}
    for adj in (0, 1):
        for order in (0, 1):
            rng = range(len(ops)) if adj == 0 else range(len(ops) - 1, -1, -1)
            shown = 0
            for oi in rng:
                a, b = snaps[0][(oi, adj, order)], snaps[1][(oi, adj, order)]
                d = np.abs(a.astype(np.float64) - b)
                rel = np.linalg.norm(d) / max(np.linalg.norm(a), 1e-30)
                if rel > 2e-5 and shown < 4:
                    shown += 1
                    idx = np.unravel_index(np.argmax(d), d.shape)
                    bad = d > 1e-4 * np.abs(a).max()
                    print("%s order %d op %3d %-8s %-28s shape %s rel %.2e  max|d| %.3e at %s (ref %.4e)  bad elems %d" % (
                        "adjoint" if adj else "value  ", order, oi, names.get(ops[oi].kind, "?"), ops[oi].name,
                        a.shape, rel, d.max(), idx, a[idx], int(bad.sum())))
                    if bad.sum():
                        w = np.argwhere(bad)
                        print("      bad n:", sorted(set(w[:, 0]))[:10], " c:", sorted(set(w[:, 1]))[:20],
                              " h:", sorted(set(w[:, 2]))[:20], " w:", sorted(set(w[:, 3]))[:20])
    _lib.check(lib.b2s_set_tensor_core_mode(1))


if __name__ == "__main__":
    main()
