"""Run a reference parameter file through the B200 path:  the equivalent of ``python main.py <pfile>``
(reference main.py:16-19) with the drop-in installed.

    python -m optwboundeigenval_b200.main <pfile> --reference /path/to/optWBoundEigenval \
        [--offline] [--set key=value ...] [--workdir DIR] [--no-install]

``--offline`` substitutes seeded synthetic data for the data modules (no network / private paths) and
builds torchvision backbones without downloading weights.  ``--set`` overrides entries of the
options dict (e.g. ``--set train=True max_iter=1 test=False``) -- several parameter files ship with
``train=False`` and only run analysis code.  ``--no-install`` runs the unmodified
reference (CPU autograd) through the same harness, for side-by-side comparison.
"""
from __future__ import annotations

import argparse
import ast
import os
import sys
import tempfile


def run(pfile, reference, offline=False, overrides=None, workdir=None, install=True, n_train=512, n_eval=128, seed=1226):
    from . import dropin
    reference = os.path.abspath(reference)
    if not os.path.exists(os.path.join(reference, "opt.py")):
        raise FileNotFoundError("no opt.py under %s" % reference)
    dropin.stub_plotting_modules()
    if reference not in sys.path:
        sys.path.insert(0, reference)
    import opt  # noqa: E402  (the reference module)
    if offline:
        dropin.offline_shims(reference, n_train=n_train, n_eval=n_eval)
    if install:
        dropin.install(opt)
    workdir = workdir or tempfile.mkdtemp(prefix="b200_run_")
    os.makedirs(workdir, exist_ok=True)
    link = os.path.join(workdir, "params")
    if not os.path.exists(link):
        os.symlink(os.path.join(reference, "params"), link)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        sys.path.insert(0, "./params")
        if seed is not None:
            # the parameter files carry 'seed': 1226 but nothing reads it: weights, the shuffling of
            # the loaders and iter()'s random end-of-epoch minibatch (opt.py:604) are seeded here so that runs repeat
            import random
            import numpy as np
            import torch
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
        params = __import__(pfile)
        if overrides:
            orig = params.options

            def options():
                d = orig()
                d.update(overrides)
                return d
            params.options = options
        opt.main(pfile)
    finally:
        os.chdir(cwd)
        if install:
            dropin.uninstall(opt)
    return workdir


def _parse_overrides(items):
    out = {}
    for it in items or []:
        k, v = it.split("=", 1)
        try:
            out[k] = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            out[k] = v
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("pfile")
    from . import dropin
    ap.add_argument("--reference", default=dropin.find_reference() or ".")
    ap.add_argument("--offline", action="store_true")
    ap.add_argument("--set", nargs="*", default=[])
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--no-install", action="store_true")
    ap.add_argument("--n-train", type=int, default=512, help="--offline: synthetic training samples")
    ap.add_argument("--n-eval", type=int, default=128, help="--offline: synthetic validation / test samples")
    ap.add_argument("--seed", type=int, default=1226, help="seed of python / numpy / torch generators (-1: leave unseeded)")
    a = ap.parse_args(argv)
    wd = run(a.pfile, a.reference, a.offline, _parse_overrides(a.set), a.workdir, not a.no_install, a.n_train, a.n_eval,
             None if a.seed < 0 else a.seed)
    print("logs and models under", wd)


if __name__ == "__main__":
    main()
