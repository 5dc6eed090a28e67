"""Where one regularised minibatch step (iter()'s body: base pass, k power iterations, vGHv, fused update) spends its
time on one GPU -- stage by stage with a device synchronisation after each (run under gpurun)."""
import contextlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import zoo                                          # noqa: E402
from optwboundeigenval_b200.spectral import SpectralState                       # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else zoo.CONFIGS[kind][3]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
model, loss = zoo.build(kind)
model = model.cuda().train()
x, y = zoo.synthetic_batch(kind, batch)
xh, yh = x.pin_memory(), y.pin_memory()
st = SpectralState(model, loss, mu=0.01, K=0.0, pow_iter_eps=0.0, max_pow_iter=k, ignore_bad_vals=False)
opt = torch.optim.SGD(model.parameters(), lr=1e-4)


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


acc = {}
with contextlib.redirect_stdout(sys.stderr):
    for it in range(8):
        t0 = sync()
        data = [xh.to("cuda", non_blocking=True), yh.to("cuda", non_blocking=True)]
        t1 = sync()
        st.comp_g(data)                       # new operator, base pass, k power iterations
        t2 = sync()
        st.fused_step(optimizer=opt)          # vGHv (order-2 pass + BatchNorm compatibility sweep), clip, assembly, update
        t3 = sync()
        if it >= 3:
            for name, dt in (("h2d", t1 - t0), ("comp_g", t2 - t1), ("fused_step", t3 - t2), ("total", t3 - t0)):
                acc.setdefault(name, []).append(dt * 1e3)
for name, v in acc.items():
    print("%-12s %8.3f ms" % (name, sum(v) / len(v)))
# the pieces of comp_g and fused_step on their own
op = st.hvp_op
t0 = sync()
for _ in range(5):
    op.prepare_grad()
t1 = sync()
print("%-12s %8.3f ms" % ("base pass", (t1 - t0) * 1e3 / 5))
v = st.v
t0 = sync()
for _ in range(5):
    op.vGHv(v, storedGrad=True)
t1 = sync()
print("%-12s %8.3f ms" % ("vGHv", (t1 - t0) * 1e3 / 5))
t0 = sync()
for _ in range(5):
    op.Hv(v, storedGrad=True)
t1 = sync()
print("%-12s %8.3f ms" % ("Hv (host api)", (t1 - t0) * 1e3 / 5))
