"""Per-kernel-family timing of the three passes for one config (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import zoo                                      # noqa: E402
from optwboundeigenval_b200.hvp_operator import B200HVPOperator             # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else zoo.CONFIGS[kind][3]
    model, loss = zoo.build(kind)
    model.train()
    x, y = zoo.synthetic_batch(kind, batch)
    op = B200HVPOperator(model, [x, y], loss)
    P = sum(p.numel() for p in model.parameters())
    v = torch.from_numpy(np.ones(P) / np.sqrt(P)).cuda()
    op.Hv(v, storedGrad=True)
    op.vGHv(v, storedGrad=True)
    torch.cuda.synchronize()
    print("config %s batch %d P %d workspace %.1f MB" % (kind, batch, P, op.plan.workspace_bytes() / 1e6))
    for order, label in ((0, "base"), (1, "Hv"), (2, "2nd order"), (3, "BN compat sweep")):
        try:
            rows = op.plan.profile(order, reps=3)
        except RuntimeError as e:
            print(label, "skipped:", e)
            continue
        tot = sum(r["ms"] for r in rows)
        print("--- pass %s: %.3f ms in kernels, %d launches" % (label, tot, sum(r["launches"] for r in rows)))
        for r in sorted(rows, key=lambda r: -r["ms"]):
            gf = r["flops"] / r["ms"] / 1e9 if r["ms"] > 0 else 0   # TFLOP/s
            gb = r["bytes"] / r["ms"] / 1e6 if r["ms"] > 0 else 0
            print("   %-16s launches %4d  %8.3f ms (%5.1f%%)  %9.2f TFLOP/s  %8.1f GB/s" % (
                r["name"], r["launches"], r["ms"], 100 * r["ms"] / tot, gf, gb))
    if os.environ.get("PROF_RAW", "0") == "1":
        rows = op.plan.profile(1, reps=3, raw=True)
        tape = op.plan.tape
        print("--- raw per-launch timeline of the Hv pass (us)")
        for r in rows:
            print("   %-24s %8.1f us  %8.2f TFLOP/s %8.1f GB/s" % (r["name"], r["ms"] * 1e3,
                  r["flops"] / r["ms"] / 1e9 if r["ms"] > 0 else 0, r["bytes"] / r["ms"] / 1e6 if r["ms"] > 0 else 0))
    # wall clock of the graph-replayed HVP
    for _ in range(3):
        op.plan.hv(v)
    torch.cuda.synchronize()
    t0 = time.time()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        op.plan.hv(v)
    e1.record()
    torch.cuda.synchronize()
    print("graph-replayed Hv: %.3f ms/HVP (events), %.3f ms wall" % (e0.elapsed_time(e1) / n, (time.time() - t0) * 1e3 / n))
    out = op.power_iterate(v, 0.0, 20)
    torch.cuda.synchronize()
    t0 = time.time()
    out = op.power_iterate(v, 0.0, 50)
    torch.cuda.synchronize()
    print("power_iterate 50 its: %.3f ms/iter wall, lam=%.6g" % ((time.time() - t0) * 1e3 / 50, out.lam))


if __name__ == "__main__":
    main()
