import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optwboundeigenval_b200 import zoo
from optwboundeigenval_b200.hvp_operator import B200HVPOperator
kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else zoo.CONFIGS[kind][3]
model, loss = zoo.build(kind); model.train()
x, y = zoo.synthetic_batch(kind, batch)
op = B200HVPOperator(model, [x, y], loss)
P = sum(p.numel() for p in model.parameters())
v = torch.from_numpy(np.ones(P) / np.sqrt(P)).cuda()
op.Hv(v, storedGrad=True)
for _ in range(5): op.plan.hv(v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): op.plan.hv(v)
e1.record(); torch.cuda.synchronize()
print(kind, batch, "HVP ms", e0.elapsed_time(e1) / 30)
