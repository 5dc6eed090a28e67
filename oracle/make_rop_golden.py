"""Fixture of the reference's closed-form R-operator specification (rop.py) -- TEST INFRASTRUCTURE ONLY.

``/root/reference/rop.py`` states, for a sigmoid MLP with square layers and the loss 0.5*||yhat - y||^2 on one
sample, the quantities of the hot path in closed form: dE/dw (rop.py:103-125), R{dE/dw} = H v (rop.py:127-143) and
R^2{dE/dw} = grad_w(v' H v) with v held fixed (rop.py:145-164) -- the specification ``HVPOperator.Hv`` / ``vGHv``
(opt.py:77-152) realise with nested autograd.  The reference ships it with MATLAB data files that are not in the
checkout, so this script writes a seeded ``.mat`` of the layout ``ROp.__init__`` reads (rop.py:44-50), runs the
UNMODIFIED class on it (build container only, where /root/reference exists) and stores inputs and outputs in
``tests/golden/rop_sigmoid_mlp.npz``.  ``tests/test_oracle_golden.py`` checks ``oracle/autograd_oracle.py``
against it in fp64, which pins the oracle's Hv / vGHv to the reference's own written specification.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_access import find_reference  # noqa: E402


def main(n: int = 7, layers: int = 3, seed: int = 20):
    import scipy.io
    ref = find_reference()
    if ref is None:
        raise FileNotFoundError("no reference checkout")
    sys.path.insert(0, ref)
    import rop  # the unmodified reference module

    rng = np.random.default_rng(seed)
    w = np.empty((layers, 1), dtype=object)
    b = np.empty((layers, 1), dtype=object)
    for i in range(layers):
        w[i, 0] = rng.standard_normal((n, n)) / np.sqrt(n) * 1.5
        b[i, 0] = 0.3 * rng.standard_normal((n, 1))
    x = rng.standard_normal((n, 1))
    y = rng.uniform(0.1, 0.9, (n, 1))
    v = rng.standard_normal((layers * n * n, 1))
    v /= np.linalg.norm(v)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "rop_case.mat")
        scipy.io.savemat(path, {"x": x, "y": y, "w": w, "b": b, "v": v})
        r = rop.ROp(path)
        r.compute()
    out = {"x": x, "y": y, "v": v, "n": np.array(n), "layers": np.array(layers)}
    for i in range(layers):
        out["w%d" % i] = w[i, 0]
        out["b%d" % i] = b[i, 0]
        out["g%d" % i] = r.dEdw[i]        # [out, in], as nn.Linear.weight
        out["hv%d" % i] = r.rdEdw[i]
        out["vghv%d" % i] = r.r2dEdw[i]
    dst = os.path.join(ROOT, "tests", "golden", "rop_sigmoid_mlp.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
