"""torchrun --nproc-per-node N tools/mgpu_check.py : sharded minibatch (N ranks, NCCL) vs the same global
batch on one GPU.  Prints relative differences of grad / Hv / vGHv and lambda_max."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import zoo                                              # noqa: E402
from optwboundeigenval_b200.hvp_operator import SpectralPlan, flat_parameters       # noqa: E402


def rel(a, b):
    return float((a - b).norm() / b.norm())


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    for kind, per_rank in (("usps", 16), ("cifar_densenet", 8), ("chest_densenet_tiny", 4)):
        if kind == "chest_densenet_tiny":
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from test_gpu_parity import _tiny_chest
            model, loss = _tiny_chest("dense"), zoo.WeightedBCEWithLogits()
            g = torch.Generator().manual_seed(9)
            X = torch.randn(per_rank * world, 3, 16, 16, generator=g)
            Y = (torch.rand(per_rank * world, 5, generator=g) > 0.7).float()
            shape = (3, 16, 16)
        else:
            model, loss = zoo.build(kind)
            X, Y = zoo.synthetic_batch(kind, per_rank * world)
            shape = zoo.CONFIGS[kind][1]
        model = model.to(dev).train()
        P = sum(p.numel() for p in model.parameters())
        gen = torch.Generator().manual_seed(3)
        v = torch.randn(P, generator=gen, dtype=torch.float64)
        v = (v / v.norm()).to(dev)
        params = flat_parameters(model)
        xs, ys = X[rank * per_rank:(rank + 1) * per_rank], Y[rank * per_rank:(rank + 1) * per_rank]
        # detach BN buffers from the comparison: every plan updates them
        sharded = SpectralPlan(model, loss, shape, per_rank, dev)
        sharded.init_comm()
        g1, l1 = sharded.base_pass(params, xs, ys)
        h1 = sharded.hv(v)
        q1 = sharded.vghv(v)
        out1 = sharded.power_iterate(v, 0.0, 5)
        torch.cuda.synchronize()
        if rank == 0:
            full = SpectralPlan(model, loss, shape, per_rank * world, dev)
            g0, l0 = full.base_pass(params, X, Y)
            h0 = full.hv(v)
            q0 = full.vghv(v)
            out0 = full.power_iterate(v, 0.0, 5)
            torch.cuda.synchronize()
            print("%-20s world=%d  loss %.8f / %.8f  grad %.2e  hv %.2e  vghv %.2e  lam %.8g / %.8g" % (
                kind, world, float(l1), float(l0), rel(g1, g0), rel(h1, h0), rel(q1, q0), out1.lam, out0.lam), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
