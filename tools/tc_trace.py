"""Debug aid (run under gpurun with B2S_TC_TRACE=1 B2S_NO_GRAPHS=1 [B2S_TC_MODE=2]): the first launches of the
tcgen05 contraction dump a per-role clock timeline of CTA 0 to stderr."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import zoo                                          # noqa: E402
from optwboundeigenval_b200.hvp_operator import B200HVPOperator                 # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cifar_densenet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else zoo.CONFIGS[kind][3]
model, loss = zoo.build(kind)
model.train()
x, y = zoo.synthetic_batch(kind, batch)
op = B200HVPOperator(model, [x, y], loss)
op.prepare_grad()
torch.cuda.synchronize()
