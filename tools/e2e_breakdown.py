"""Where the end-to-end HVP (host vector in, host vector out) spends its time beyond the device-resident step."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import zoo                                      # noqa: E402
from optwboundeigenval_b200.hvp_operator import B200HVPOperator             # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
batch = zoo.CONFIGS[kind][3]
model, loss = zoo.build(kind)
model.train()
x, y = zoo.synthetic_batch(kind, batch)
op = B200HVPOperator(model, [x, y], loss)
op.async_host_vectors = True
P = sum(p.numel() for p in model.parameters())
v = torch.from_numpy(np.ones(P) / np.sqrt(P)).cuda()
op.Hv(v, storedGrad=True)
N = 50


def timed(fn, n=N, sync_each=False):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


out = op.power_iterate(v, 0.0, 5)
torch.cuda.synchronize()
t0 = time.perf_counter()
op.power_iterate(v, 0.0, N)
torch.cuda.synchronize()
print("power_iterate per iteration      %.3f ms" % ((time.perf_counter() - t0) / N * 1e3))
print("plan.hv(device v), queued        %.3f ms" % timed(lambda: op.plan.hv(v)))
print("plan.hv(device v), sync each     %.3f ms" % timed(lambda: op.plan.hv(v), sync_each=True))
print("op.Hv(device v), sync each       %.3f ms" % timed(lambda: op.Hv(v, storedGrad=True), sync_each=True))
vh = torch.from_numpy(np.ones(P) / np.sqrt(P)).pin_memory()
rh = torch.empty(P, dtype=torch.float64).pin_memory()
cur = torch.cuda.current_stream()


def e2e():
    r = op.Hv(vh, storedGrad=True)
    rh.copy_(r, non_blocking=True)
    cur.synchronize()


print("op.Hv(pinned host v) + D2H + sync %.3f ms" % timed(e2e))
a, b = vh.numpy(), rh.numpy()


def full():
    e2e()
    np.multiply(b, 1.0 / np.sqrt(np.dot(b, b)), out=a)


print("... + host normalisation         %.3f ms" % timed(full))
t0 = time.perf_counter()
for _ in range(200):
    np.multiply(b, 1.0 / np.sqrt(np.dot(b, b)), out=a)
print("host normalisation alone         %.3f ms" % ((time.perf_counter() - t0) / 200 * 1e3))
d = torch.empty(P, dtype=torch.float64, device="cuda")


def copies():
    d.copy_(vh, non_blocking=True)
    rh.copy_(d, non_blocking=True)
    cur.synchronize()


print("H2D + D2H of 8P bytes + sync     %.3f ms" % timed(copies))
t0 = time.perf_counter()
for _ in range(200):
    op._ensure_grad(True)
    op.plan._bind_stream()
    torch.empty(P, dtype=torch.float64, device="cuda")
print("python-side bookkeeping per call %.3f ms" % ((time.perf_counter() - t0) / 200 * 1e3))
