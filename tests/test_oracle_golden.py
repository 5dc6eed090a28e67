"""Pins oracle/autograd_oracle.py against vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, model_from_golden, rel_err
from oracle import autograd_oracle as ao

CASES = [("forest", "forest"), ("usps", "usps"), ("cifar_densenet", "cifar_densenet")]


@pytest.mark.parametrize("name,kind", CASES)
def test_grad_hv_vghv_match_reference(name, kind):
    g = load_golden(name)
    model, loss = model_from_golden(kind, g)
    op = ao.AutogradSpectralOperator(model, [torch.from_numpy(g["x"]), torch.from_numpy(g["y"])], loss)
    P = g["grad"].size
    assert rel_err(op.gradient().detach().numpy(), g["grad"]) < 2e-6
    assert rel_err(op.hv(ao.start_vector(P)).numpy(), g["hv_v0"]) < 2e-5
    vr = torch.from_numpy(g["v_rand"])
    assert rel_err(op.hv(vr).numpy(), g["hv_vrand"]) < 2e-5
    assert rel_err(op.vghv(vr).numpy(), g["vghv_vrand"]) < 2e-5
    assert abs(op.loss_value - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))


@pytest.mark.parametrize("name,kind", CASES)
def test_power_iteration_matches_reference(name, kind):
    g = load_golden(name)
    meta = eval(str(g["meta"]))
    model, loss = model_from_golden(kind, g)
    op = ao.AutogradSpectralOperator(model, [torch.from_numpy(g["x"]), torch.from_numpy(g["y"])], loss)
    P = g["grad"].size
    res = ao.power_iteration(op.hv, ao.start_vector(P), eps=meta["eps"], max_iter=meta["max_pow_iter"])
    assert res["iters"] == int(g["rho1_iters"])
    assert abs(res["rho"] - float(g["rho1_rho"])) <= 1e-3 * float(g["rho1_rho"])
    v_ref = g["rho1_v"]
    v = res["v"].numpy()
    assert min(rel_err(v, v_ref), rel_err(-v, v_ref)) < 1e-3
    traj = np.array(res["trajectory"])
    np.testing.assert_allclose(traj[:, 1], g["rho1_traj"][:, 1], rtol=1e-3)
    gr = ao.penalty_gradient(op, res["v"])
    assert rel_err(gr.numpy(), g["rho1_gradrho"]) < 1e-3
    if "rho2_iters" in g:   # warm start on a second batch
        op2 = ao.AutogradSpectralOperator(model, [torch.from_numpy(g["x2"]), torch.from_numpy(g["y2"])], loss)
        res2 = ao.power_iteration(op2.hv, res["v"], eps=meta["eps"], max_iter=meta["max_pow_iter"])
        assert res2["iters"] == int(g["rho2_iters"])
        assert abs(res2["rho"] - float(g["rho2_rho"])) <= 1e-3 * float(g["rho2_rho"])


@pytest.mark.parametrize("name,kind,alpha", [
    ("forest_lobpcg", "forest", lambda k: np.exp(-4 * k - 2)),
    ("usps_lobpcg", "usps", lambda k: np.exp(-4 * k)),
])
def test_kfac_preconditioned_iteration_matches_reference(name, kind, alpha):
    g = load_golden(name)
    meta = eval(str(g["meta"]))
    model, loss = model_from_golden(kind, g)
    data = [torch.from_numpy(g["x"]), torch.from_numpy(g["y"])]
    pre = ao.KfacPreconditioner(model)
    pre.build(data, loss)
    op = ao.AutogradSpectralOperator(model, data, loss)
    P = g["grad"].size
    res = ao.power_iteration(op.hv, ao.start_vector(P), eps=meta["eps"], max_iter=meta["max_pow_iter"],
                             alpha=alpha, precond=pre.apply)
    assert res["iters"] == int(g["rho1_iters"])
    assert abs(res["rho"] - float(g["rho1_rho"])) <= 1e-3 * float(g["rho1_rho"])
    v = res["v"].numpy()
    assert min(rel_err(v, g["rho1_v"]), rel_err(-v, g["rho1_v"])) < 1e-3


def test_bn_running_stats_side_effect():
    g = load_golden("cifar_densenet")
    model, loss = model_from_golden("cifar_densenet", g)
    op = ao.AutogradSpectralOperator(model, [torch.from_numpy(g["x"]), torch.from_numpy(g["y"])], loss)
    op.gradient()
    flat = np.concatenate([v.detach().reshape(-1).double().numpy() for v in model.state_dict().values()])
    assert rel_err(flat, g["state_after_one_pass"]) < 1e-6


def test_closed_form_rop_specification():
    """rop.py (the reference's written R-operator specification: sigmoid MLP, 0.5*||yhat - y||^2, one sample) run
    unmodified by oracle/make_rop_golden.py: dE/dw, R{dE/dw} = Hv and R^2{dE/dw} = vGHv (rop.py:103-164) must be what
    the autograd oracle computes for the same network, weights and direction (fp64, biases carry no direction)."""
    g = load_golden("rop_sigmoid_mlp")
    n, L = int(g["n"]), int(g["layers"])
    layers = []
    for i in range(L):
        lin = torch.nn.Linear(n, n).double()
        lin.weight.data = torch.from_numpy(g["w%d" % i]).clone()
        lin.bias.data = torch.from_numpy(g["b%d" % i][:, 0]).clone()
        layers += [lin, torch.nn.Sigmoid()]
    model = torch.nn.Sequential(*layers)
    x = torch.from_numpy(g["x"].T.copy())
    y = torch.from_numpy(g["y"].T.copy())
    op = ao.AutogradSpectralOperator(model, [x, y], lambda out, tgt: 0.5 * ((out - tgt) ** 2).sum())
    # direction: rop.py:81 reshapes slice i column-major into the [out, in] matrix of layer i; nn.Linear is row-major
    parts = []
    for i in range(L):
        parts += [np.reshape(g["v"][i * n * n:(i + 1) * n * n, 0], (n, n), order="F").reshape(-1), np.zeros(n)]
    v = torch.from_numpy(np.concatenate(parts))
    grad, hv, vghv = op.gradient().detach().numpy(), op.hv(v).numpy(), op.vghv(v).numpy()
    off = 0
    for i in range(L):
        sl = slice(off, off + n * n)
        assert rel_err(grad[sl], g["g%d" % i].reshape(-1)) < 1e-12
        assert rel_err(hv[sl], g["hv%d" % i].reshape(-1)) < 1e-12
        assert rel_err(vghv[sl], g["vghv%d" % i].reshape(-1)) < 1e-11
        off += n * n + n
