"""Short single-GPU program for ncu: base pass + a few HVPs of one config (graphs off so every kernel
is a separate, named launch)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import _lib, zoo                                # noqa: E402
from optwboundeigenval_b200.hvp_operator import B200HVPOperator             # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else zoo.CONFIGS[kind][3]
n_hv = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model, loss = zoo.build(kind)
model.train()
x, y = zoo.synthetic_batch(kind, batch)
op = B200HVPOperator(model, [x, y], loss)
P = sum(p.numel() for p in model.parameters())
v = torch.from_numpy(np.ones(P) / np.sqrt(P)).cuda()
op.prepare_grad()
_lib.check(op.plan.lib.b2s_plan_set_graphs(op.plan.handle, 0))
op.stored_grad = op.prepare_grad()
for _ in range(n_hv):
    hv = op.Hv(v, storedGrad=True)
torch.cuda.synchronize()
print("ok", float(hv.norm()))
