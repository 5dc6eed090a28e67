"""The five BASELINE parameter files through the reference's own ``opt.main`` with the drop-in installed
(``python -m optwboundeigenval_b200.main <pfile> --offline``), on a GPU, against the same run of the UNMODIFIED
reference on the CPU (``--no-install``): the epoch log line ``epoch f rho h norm`` (opt.py:800-832) must agree.

The reference checkout is found under ``baseline/_ref/optWBoundEigenval`` on the GPU box (``dropin.find_reference``).
Synthetic data (``--offline``), one epoch, few minibatches: several files ship with ``train=False``, so
the harness switches training on, exactly as a user reproducing the experiment would.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200.dropin import find_reference  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(find_reference() is None, reason="no reference checkout (baseline/_ref/optWBoundEigenval)")]

COMMON = ["train=True", "test=False", "max_iter=1", "comp_test=False", "rho_test=False", "jaccard=False",
          "jaccard_comp=False", "saliency=0", "crops=False", "num_workers=0"]
CASES = {
    # pfile: (extra overrides, n_train, n_eval, rtol on rho)
    "forest_best": ([], 256, 64, 2e-3),
    "usps_CNN_mu0_01_K0": ([], 256, 64, 2e-3),
    "usps_CNN_lobpcg": ([], 256, 64, 2e-3),
    "cifar10_DenseNet_mu0_01_K10": (["max_pow_iter=6"], 64, 32, 1e-2),
    # 32 validation samples: the AUC of test_model needs both labels in each of the 14 classes, or no "best" model is saved
    "chestxray_best_reg": (["max_pow_iter=2"], 4, 32, 2e-2),
}


def _run(pfile, workdir, install, extra, n_train, n_eval):
    cmd = [sys.executable, "-m", "optwboundeigenval_b200.main", pfile, "--offline", "--workdir", str(workdir),
           "--n-train", str(n_train), "--n-eval", str(n_eval), "--set"] + COMMON + extra
    if not install:
        cmd += ["use_gpu=False", "--no-install"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, "%s (install=%s) failed:\n%s" % (pfile, install, (r.stdout + r.stderr)[-3000:])
    logs = [f for f in os.listdir(os.path.join(workdir, "logs")) if f.endswith(".log") and "verbose" not in f]
    assert len(logs) == 1, logs
    rows = [ln.split("\t") for ln in open(os.path.join(workdir, "logs", logs[0])).read().splitlines()
            if ln and ln[0].isdigit()]
    assert rows, "no epoch line in %s" % logs[0]
    return [float(t) for t in rows[0]]          # epoch, f, rho, h, norm[, val_acc, val_f1]


@pytest.mark.parametrize("pfile", list(CASES))
def test_parameter_file_runs_through_the_dropin_and_matches_the_reference(pfile, tmp_path):
    extra, n_train, n_eval, rtol = CASES[pfile]
    mine = _run(pfile, tmp_path / "b200", True, extra, n_train, n_eval)
    ref = _run(pfile, tmp_path / "ref", False, extra, n_train, n_eval)
    print(pfile, "b200", mine, "reference", ref)
    assert len(mine) == len(ref)
    f, rho, h = mine[1], mine[2], mine[3]
    assert abs(f - ref[1]) <= 2e-3 * abs(ref[1])
    if ref[2] > 0:
        assert abs(rho - ref[2]) <= rtol * ref[2]
    else:
        assert rho == ref[2]                   # the -1 sentinel of a non-converged run
    assert abs(h - ref[3]) <= max(2e-3 * abs(ref[3]), rtol * abs(ref[3]))
    if len(ref) > 5:                            # validation accuracy / F1 through test_model -> comp_f
        assert abs(mine[5] - ref[5]) <= 2.0     # percent (or AUC): a handful of borderline predictions at most
