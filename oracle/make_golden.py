"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):   python oracle/make_golden.py [--out tests/golden] [--only forest,usps,...]

Recipe: ``import pandas`` first, stub ``matplotlib``,
``matplotlib.pyplot`` and ``pytz`` (imported at module scope by opt.py:17,34 and
dcnn.py:9), put the reference on ``sys.path``, build the reference's own model
classes, feed seeded synthetic batches, and record what ``opt.HVPOperator`` /
``opt.OptWBoundEignVal.comp_rho`` / ``comp_gradrho`` return.  The stored arrays
are what the reference computed -- nothing from this repo is on that path except
the synthetic input generator.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("OPTW_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


sys.path.insert(0, ROOT)
from oracle.reference_access import import_reference as _import_reference, ref_model  # noqa: E402


def import_reference():
    return _import_reference(REF if os.path.exists(os.path.join(REF, "opt.py")) else None)


def flat_state(model):
    import torch
    names, arrs = [], []
    for k, v in model.state_dict().items():
        names.append(k)
        arrs.append(v.detach().reshape(-1).double().numpy())
    return names, np.concatenate(arrs).astype(np.float32)


SAMPLE = 1 << 17          # fixed-index sample stored for the parameter-sized vectors of the chest models


def sample_index(n: int) -> np.ndarray:
    """The fixed index set of the compact fixtures (tests/conftest.py restates it)."""
    if n <= SAMPLE:
        return np.arange(n)
    return np.sort(np.random.RandomState(20261018).choice(n, SAMPLE, replace=False))


def checksum(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    w = np.cos(np.arange(a.size, dtype=np.float64) * 0.61803398875)      # position-sensitive
    return np.array([a.sum(), np.square(a).sum(), np.dot(a, w)])


class _Packer:
    """Full vectors for the small models; for the chest models (P = 7e6 / 1.9e7) the norm, a fixed 2^17-index
    sample and the classifier tail -- a 78 MB vector per entry is not a fixture."""

    def __init__(self, out, compact, n, tail):
        self.out, self.compact, self.tail = out, compact, tail
        self.idx = sample_index(n) if compact else None

    def put(self, key, vec, dtype=np.float32):
        vec = np.asarray(vec)
        if not self.compact:
            self.out[key] = vec.astype(dtype)
            return
        self.out[key + "_norm"] = np.array(float(np.linalg.norm(vec.astype(np.float64))))
        self.out[key + "_s"] = vec[self.idx].astype(dtype)
        self.out[key + "_tail"] = vec[-self.tail:].astype(dtype)


def run_case(opt, kind: str, batch: int, *, eps: float, max_pow_iter: int, lobpcg: bool = False,
             alpha=1, kfac_batch: int = 1, clip=None, second_batch: bool = True, seed: int = 1226,
             compact: bool = False, rand_init: bool = False, ignore_bad_vals: bool = False, Kmin: float = 0,
             K: float = 0, kfac_rand: bool = False):
    import torch
    sys.path.insert(0, ROOT)
    from optwboundeigenval_b200 import zoo

    torch.manual_seed(seed)
    np.random.seed(seed)
    model, loss = ref_model(kind)
    model.train()
    names, state0 = flat_state(model)
    x, y = zoo.synthetic_batch(kind, batch, seed)
    x2, y2 = zoo.synthetic_batch(kind, batch, seed + 7)
    ndim = sum(p.numel() for p in model.parameters())
    out = {"state_names": np.array(names), "y": y.numpy()}
    if compact:
        # the zoo model built with the same seed must carry the reference's weights (checked here and by the test)
        zm, _ = zoo.build(kind, seed)
        assert np.array_equal(flat_state(zm)[1], state0), "zoo.build(%s) does not reproduce the reference's weights" % kind
        out["state0_check"] = checksum(state0)
        out["x_check"] = checksum(x.numpy())
    else:
        out.update({"state0": state0, "x": x.numpy()})
        if second_batch:
            out.update({"x2": x2.numpy(), "y2": y2.numpy()})
    tail = sum(p.numel() for n_, p in model.named_parameters() if "classifier" in n_) if compact else 0
    pk = _Packer(out, compact, ndim, tail)
    v0 = torch.from_numpy(np.ones(ndim) / np.sqrt(ndim))
    gen = torch.Generator().manual_seed(seed + 3)
    vr = torch.randn(ndim, generator=gen, dtype=torch.float64)
    vr /= vr.norm()
    if compact:
        out["v_rand_check"] = checksum(vr.numpy())
    else:
        out["v_rand"] = vr.numpy()

    # --- HVPOperator on batch 1 (fresh operator per vGHv: it is one-shot)
    op = opt.HVPOperator(model, [x, y], loss, use_gpu=False)
    pk.put("hv_v0", op.Hv(v0, storedGrad=True).numpy())
    pk.put("grad", op.stored_grad.detach().numpy())
    pk.put("hv_vrand", op.Hv(vr.numpy(), storedGrad=True).numpy())   # ndarray input path
    pk.put("vghv_vrand", op.vGHv(vr, storedGrad=True).numpy())
    # module buffers after ONE train-mode forward (BN running stats side effect)
    model2, _ = ref_model(kind)
    model2.load_state_dict(dict(zip(names, _unflatten(state0, model2))))
    model2.train()
    op2 = opt.HVPOperator(model2, [x, y], loss, use_gpu=False)
    op2.Hv(v0, storedGrad=True)
    if compact:                                  # only the buffers change: keep those
        sd2 = model2.state_dict()
        pnames = {n_ for n_, _ in model2.named_parameters()}
        out["buffers_after_one_pass"] = np.concatenate(
            [sd2[k].detach().reshape(-1).double().numpy() for k in names if k not in pnames]).astype(np.float32)
    else:
        out["state_after_one_pass"] = flat_state(model2)[1]
    out["loss"] = np.array(_train_loss(kind, names, state0, x, y))

    # --- comp_rho / comp_gradrho through OptWBoundEignVal on a pristine copy of the weights
    model3, loss3 = ref_model(kind)
    model3.load_state_dict(dict(zip(names, _unflatten(state0, model3))))
    calls = _Trajectory()

    class Recording(opt.HVPOperator):           # records, does not alter, the reference operator
        def Hv(self, vec, storedGrad=False):
            r = super().Hv(vec, storedGrad)
            calls.add(torch.as_tensor(vec).detach().double(), r)
            return r

    real = opt.HVPOperator
    opt.HVPOperator = Recording
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "logs"))
    os.chdir(tmp)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            o = opt.OptWBoundEignVal(model3, loss3, torch.optim.SGD(model3.parameters(), lr=0.1), mu=0.01, K=0,
                                     pow_iter_eps=eps, max_pow_iter=max_pow_iter, use_gpu=False,
                                     ignore_bad_vals=ignore_bad_vals, verbose=True, header="golden", lobpcg=lobpcg,
                                     pow_iter_alpha=alpha, kfac_batch=kfac_batch, kfac_rand=kfac_rand,
                                     gradg_clip=clip, batch_size=batch, rand_init=rand_init, Kmin=Kmin)
            o.K = float(K)
            if kfac_rand:
                torch.manual_seed(seed + 11)       # opt.py:351-356 samples the targets from the global CPU generator
            i, rn, size = o.comp_rho([x, y])
        out["rho1_iters"] = np.array(i)
        out["rho1_rn"] = np.array(float(rn))
        out["rho1_rho"] = np.array(float(o.rho))
        out["rho1_norm"] = np.array(float(o.norm))
        pk.put("rho1_v", o.v.numpy().copy(), np.float64)
        out["rho1_traj"] = calls.rows()
        with contextlib.redirect_stdout(io.StringIO()):
            o.g = np.max([0.0, o.rho - o.K, o.Kmin - o.rho])     # opt.py:578
            o.comp_gradrho()
        out["rho1_g"] = np.array(float(o.g))
        out["rho1_gradrho_norm"] = np.array(float(torch.norm(o.gradrho)))      # after the clip of opt.py:539-542
        pk.put("rho1_gradrho", o.gradrho.numpy())
        pk.put("rho1_gradf", o.hvp_op.stored_grad.detach().numpy())
        if second_batch:                        # warm start from batch 1's vector (opt.py:432)
            calls.clear()
            with contextlib.redirect_stdout(io.StringIO()):
                i2, rn2, _ = o.comp_rho([x2, y2])
            out["rho2_iters"] = np.array(i2)
            out["rho2_rho"] = np.array(float(o.rho))
            out["rho2_norm"] = np.array(float(o.norm))
            pk.put("rho2_v", o.v.numpy().copy(), np.float64)
            out["rho2_traj"] = calls.rows()
        with open(os.path.join(tmp, "logs", os.path.basename(o.verbose_log_file))) as fh:
            out["verbose_log"] = np.array(fh.read())
    finally:
        os.chdir(cwd)
        opt.HVPOperator = real
    return out


def _train_loss(kind, names, state0, x, y):
    import torch
    m, l = ref_model(kind)
    m.load_state_dict(dict(zip(names, _unflatten(state0, m))))
    m.train()
    with torch.no_grad():
        return float(l(m(x), y))


def _unflatten(flat, model):
    import torch
    res, j = [], 0
    for k, v in model.state_dict().items():
        n = v.numel()
        res.append(torch.from_numpy(np.asarray(flat[j:j + n])).to(v.dtype).view(v.shape))
        j += n
    return res


class _Trajectory:
    """(i, lam, n, rn, vnn) per iteration, recomputed in fp64 from the (v, Hv) pairs the reference operator saw,
    exactly as opt.py:455-464 does; only the previous residual is kept (a chest vector is 156 MB)."""

    def __init__(self):
        self.clear()

    def clear(self):
        self._rows, self._r_old = [], 0.0

    def add(self, v, w):
        import torch
        lam = float(torch.dot(w, v))
        if lam < 0:
            lam, w = -lam, -w
        r = w - lam * v
        n = float(torch.norm(r))
        rn = float(min(torch.norm(r - self._r_old), torch.norm(r + self._r_old)))
        self._rows.append((len(self._rows), lam, n, rn, float(torch.norm(w))))
        self._r_old = r

    def rows(self):
        return np.array(self._rows, dtype=np.float64)


def run_iter_case(opt, kind: str, batch: int, n_batches: int, *, optimizer: str, eps: float, max_pow_iter: int,
                  mu: float = 0.01, K: float = 0, clip=None, seed: int = 1226):
    """One epoch of the UNMODIFIED ``OptWBoundEignVal.iter()`` (opt.py:580-763) over ``n_batches`` minibatches:
    the parameters after the optimizer steps, the per-minibatch verbose line (rho, norm, |grad f|, |grad g|) and
    the epoch figures f, rho, g, h -- the fixture of the fused step (comp_g -> grad f + mu * sign * grad rho ->
    clip -> param.grad -> SGD-momentum / Adam update)."""
    import random
    import torch
    sys.path.insert(0, ROOT)
    from optwboundeigenval_b200 import zoo
    torch.manual_seed(seed)
    np.random.seed(seed)
    model, loss = ref_model(kind)
    model.train()
    names, state0 = flat_state(model)
    xs, ys = zip(*[zoo.synthetic_batch(kind, batch, seed + 100 + j) for j in range(n_batches)])
    x, y = torch.cat(xs), torch.cat(ys)
    if optimizer == "sgd":          # params/cifar10_DenseNet_mu0_01_K10.py:47
        optim = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    else:                           # params/usps_CNN_lobpcg.py:46 / chestxray_best_reg.py:106 (with weight decay)
        optim = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "logs"))
    os.chdir(tmp)
    out = {"state_names": np.array(names), "state0": state0, "x": x.numpy(), "y": y.numpy()}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            o = opt.OptWBoundEignVal(model, loss, optim, mu=mu, K=K, pow_iter_eps=eps, max_pow_iter=max_pow_iter,
                                     use_gpu=False, ignore_bad_vals=False, verbose=True, header="golden_iter",
                                     gradg_clip=clip, batch_size=batch)
            o.dataloader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=batch)
            random.seed(seed)                   # iter() draws the end-of-epoch batch with random.randint (opt.py:604)
            o.iter()
        out["state_after_iter"] = flat_state(model)[1]
        out["f"], out["rho"], out["g"], out["h"] = (np.array(float(t)) for t in (o.f, o.rho, o.g, o.h))
        with open(os.path.join(tmp, "logs", os.path.basename(o.verbose_log_file))) as fh:
            out["verbose_log"] = np.array(fh.read())
    finally:
        os.chdir(cwd)
    return out


CASES = {
    # name: kwargs  (sizes keep the fixtures a few MB in total)
    "forest": dict(kind="forest", batch=128, eps=1e-3, max_pow_iter=1000),
    "forest_lobpcg": dict(kind="forest", batch=128, eps=1e-3, max_pow_iter=1000, lobpcg=True,
                          alpha=lambda k: np.exp(-4 * k - 2), kfac_batch=2),
    # the BASELINE batch sizes: params/usps_CNN_mu0_01_K0.py:27 (128), cifar10_DenseNet_mu0_01_K10.py:27 (32),
    # chestxray_best_reg.py:26 (4)
    "usps": dict(kind="usps", batch=128, eps=1e-3, max_pow_iter=30),
    "usps_lobpcg": dict(kind="usps", batch=128, eps=1e-3, max_pow_iter=1000, lobpcg=True,
                        alpha=lambda k: np.exp(-4 * k), kfac_batch=4),
    "cifar_densenet": dict(kind="cifar_densenet", batch=32, eps=5e-2, max_pow_iter=100, second_batch=False),
    # the flags of params/chestxray_best_reg.py:110-134 (gradg_clip=100, rand_init=True, eps 0.1, 100 iterations)
    "chest_vgg": dict(kind="chest_vgg", batch=4, eps=0.1, max_pow_iter=100, second_batch=False, compact=True,
                      rand_init=True, clip=100),
    "chest_densenet121": dict(kind="chest_densenet121", batch=4, eps=0.1, max_pow_iter=100, second_batch=False,
                              compact=True, rand_init=True, clip=100),
    # flag coverage on a small model: non-converged run with ignore_bad_vals (rho = -1 sentinel, v reset,
    # opt.py:513-520), Kmin > 0 (opt.py:578), a clip that engages (opt.py:539-542), rand_init (opt.py:432)
    "usps_flags": dict(kind="usps", batch=32, eps=1e-7, max_pow_iter=6, ignore_bad_vals=True, Kmin=0.5, clip=1e-3,
                       rand_init=True),
    "usps_kmin": dict(kind="usps", batch=32, eps=1e-3, max_pow_iter=200, Kmin=5.0, K=1e9, clip=1e-3, second_batch=False),
    "usps_lobpcg_rand": dict(kind="usps", batch=32, eps=1e-3, max_pow_iter=1000, lobpcg=True, kfac_rand=True,
                             alpha=lambda k: np.exp(-4 * k), kfac_batch=1, second_batch=False),
}

ITER_CASES = {
    "usps_iter_sgd": dict(kind="usps", batch=32, n_batches=3, optimizer="sgd", eps=1e-3, max_pow_iter=200, clip=0.05),
    "usps_iter_adam": dict(kind="usps", batch=32, n_batches=3, optimizer="adam", eps=1e-3, max_pow_iter=200),
    "cifar_iter_sgd": dict(kind="cifar_densenet", batch=8, n_batches=2, optimizer="sgd", eps=5e-2, max_pow_iter=100,
                           K=0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    opt = import_reference()
    os.makedirs(args.out, exist_ok=True)
    only = [s for s in args.only.split(",") if s]
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        res = run_case(opt, **kw)
        meta = {k: v for k, v in kw.items() if not callable(v)}
        res["meta"] = np.array(repr(meta))
        path = os.path.join(args.out, name + ".npz")
        np.savez_compressed(path, **res)
        print("%-16s -> %s  (%.2f MB)  rho=%g iters=%d" % (
            name, path, os.path.getsize(path) / 1e6, float(res["rho1_rho"]), int(res["rho1_iters"])), flush=True)
    for name, kw in ITER_CASES.items():
        if only and name not in only:
            continue
        res = run_iter_case(opt, **kw)
        res["meta"] = np.array(repr(kw))
        path = os.path.join(args.out, name + ".npz")
        np.savez_compressed(path, **res)
        print("%-16s -> %s  (%.2f MB)  f=%g rho=%g g=%g" % (
            name, path, os.path.getsize(path) / 1e6, float(res["f"]), float(res["rho"]), float(res["g"])), flush=True)


if __name__ == "__main__":
    main()
