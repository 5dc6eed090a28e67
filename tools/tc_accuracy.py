"""Accuracy probe of the tcgen05 3xTF32 contraction (run under gpurun): one wide 3x3 convolution
(K = Cin*9) evaluated by the tensor-core path and by the fp32 CUDA-core path, both compared with an
fp64 convolution of the same fp32 inputs.  Reports relative L2 error, mean signed error (a biased
accumulator rounding shows up there) and max error.  Test infrastructure."""
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


class OneConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.c = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cout, 10)

    def forward(self, x):
        return self.fc(torch.flatten(self.pool(self.c(x)), 1))


def read_tensor(plan, adjoint, order, t, batch):
    vt = plan.tape.tensors[t]
    out = np.zeros((batch,) + tuple(vt.shape), dtype=np.float32)
    _lib.check(plan.lib.b2s_debug_read(plan.handle, adjoint, order, t, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def main():
    lib = _lib.load()
    for cin, cout, positive in ((512, 128, False), (512, 128, True), (64, 128, True)):
        torch.manual_seed(3)
        m = OneConv(cin, cout).train()
        x = torch.randn(4, cin, 16, 16)
        if positive:            # all-positive operands: no cancellation, rounding bias is visible
            x = x.abs()
            with torch.no_grad():
                m.c.weight.abs_()
        y = torch.randint(0, 10, (4,))
        want = torch.nn.functional.conv2d(x.double(), m.c.weight.detach().double(), padding=1).numpy()
        for mode in (0, 2):
            clear_plans()
            _lib.check(lib.b2s_set_tensor_core_mode(mode))
            op = B200HVPOperator(m, [x, y], nn.CrossEntropyLoss())
            op.prepare_grad()
            torch.cuda.synchronize()
            plan = op.plan
            conv_op = [o for o in plan.tape.ops if o.kind == 1][0]
            got = read_tensor(plan, 0, 0, conv_op.out, 4).astype(np.float64)
            d = got - want
            print("K=%5d positive=%d mode=%d  rel_l2 %.3e  mean_signed_rel %.3e  max_rel %.3e" % (
                cin * 9, positive, mode, np.linalg.norm(d) / np.linalg.norm(want),
                float(np.mean(d / np.abs(want).mean())), float(np.abs(d).max() / np.abs(want).mean())))
    _lib.check(lib.b2s_set_tensor_core_mode(1))
    clear_plans()


if __name__ == "__main__":
    main()
