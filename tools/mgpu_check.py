"""torchrun --nproc-per-node N tools/mgpu_check.py : the sharded minibatch (N ranks, synced BatchNorm sums,
all-reduced results) against the same global batch on one GPU.  Checks grad / Hv / vGHv / lambda_max, ragged
shards, the batch-global weighted-BCE counts and the K-FAC preconditioned variant under data parallelism; prints one
line per case and exits non-zero when a difference exceeds its tolerance (tests/test_gpu_multi.py runs it)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import hvp_operator, zoo                                # noqa: E402
from optwboundeigenval_b200.hvp_operator import SpectralPlan, flat_parameters       # noqa: E402

TOL = 1e-4          # north-star rtol on vectors
TOL_LAM = 1e-3


def rel(a, b):
    return float((a - b).norm() / b.norm())


def shard_bounds(total, world, ragged):
    """equal shards, or a ragged split (rank 0 gets the remainder on top)"""
    if not ragged:
        per = total // world
        return [(r * per, (r + 1) * per) for r in range(world)]
    per = total // world - 2
    cuts = [0]
    for r in range(world):
        cuts.append(cuts[-1] + (per if r else total - per * (world - 1)))
    return list(zip(cuts[:-1], cuts[1:]))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    failures = []
    cases = [("usps", 16, False), ("cifar_densenet", 8, False), ("cifar_densenet", 8, True),
             ("chest_densenet_tiny", 4, False)]
    for kind, per_rank, ragged in cases:
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
        lo, hi = shard_bounds(per_rank * world, world, ragged)[rank]
        xs, ys = X[lo:hi], Y[lo:hi]
        sharded = SpectralPlan(model, loss, shape, per_rank * world, dev)
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
            errs = (rel(g1, g0), rel(h1, h0), rel(q1, q0), abs(out1.lam - out0.lam) / abs(out0.lam),
                    abs(float(l1) - float(l0)) / abs(float(l0)))
            print("%-20s world=%d %s loss %.8f / %.8f  grad %.2e  hv %.2e  vghv %.2e  lam %.8g / %.8g" % (
                kind, world, "ragged" if ragged else "equal ", float(l1), float(l0), errs[0], errs[1], errs[2],
                out1.lam, out0.lam), flush=True)
            if max(errs[:3]) > TOL or errs[3] > TOL_LAM or errs[4] > 1e-5:
                failures.append((kind, ragged, errs))
        dist.barrier()

    # K-FAC preconditioned variant (lobpcg=True, opt.py:362-416) under data parallelism: factors are batch means
    from optwboundeigenval_b200.spectral import SpectralState
    model, loss = zoo.build("usps")
    model = model.to(dev).train()
    X, Y = zoo.synthetic_batch("usps", 16 * world)
    alpha = lambda k: float(np.exp(-4 * k))     # noqa: E731
    st = SpectralState(model, loss, pow_iter_eps=1e-3, max_pow_iter=50, ignore_bad_vals=False, lobpcg=True, kfac_batch=1,
                       kfac_rand=False, pow_iter_alpha=alpha)
    i1, _, _ = st.comp_rho([X[rank * 16:(rank + 1) * 16], Y[rank * 16:(rank + 1) * 16]])
    rho1, v1 = st.rho, st.v.clone()
    gen = torch.Generator().manual_seed(1)
    r = torch.randn(st.ndim, generator=gen, dtype=torch.float64).to(dev)
    t1 = st.kfac(r).clone()
    dist.barrier()
    if rank == 0:
        hvp_operator.set_data_parallel(False)
        st0 = SpectralState(model, loss, pow_iter_eps=1e-3, max_pow_iter=50, ignore_bad_vals=False, lobpcg=True,
                            kfac_batch=1, kfac_rand=False, pow_iter_alpha=alpha)
        i0, _, _ = st0.comp_rho([X, Y])
        t0 = st0.kfac(r)
        errs = (rel(t1, t0), abs(rho1 - st0.rho) / st0.rho, min(rel(v1, st0.v), rel(-v1, st0.v)))
        print("%-20s world=%d        iters %d / %d  T r %.2e  rho %.8g / %.8g  v %.2e" % (
            "usps_lobpcg (K-FAC)", world, i1, i0, errs[0], rho1, st0.rho, errs[2]), flush=True)
        if i1 != i0 or errs[0] > TOL or errs[1] > TOL_LAM or errs[2] > 1e-3:
            failures.append(("usps_lobpcg", False, errs))
        hvp_operator.set_data_parallel(True)
    dist.barrier()
    flag = torch.tensor([len(failures)], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if int(flag.item()):
        if rank == 0:
            print("FAILED:", failures, flush=True)
        sys.exit(1)
    if rank == 0:
        print("mgpu_check ok", flush=True)


if __name__ == "__main__":
    main()
