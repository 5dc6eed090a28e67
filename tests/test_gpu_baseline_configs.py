"""GPU parity on the BASELINE configurations at their own batch sizes and with their own flags, against fixtures
produced by the UNMODIFIED reference (oracle/make_golden.py):

* chest VGG16-bn / DenseNet121 at B = 4, 224 x 224 with ``rand_init=True, gradg_clip=100, pow_iter_eps=0.1,
  max_pow_iter=100`` (params/chestxray_best_reg.py:110-134): grad, Hv, vGHv, comp_rho, clipped penalty gradient;
* the flags no other test reaches: ``ignore_bad_vals`` (rho = -1 sentinel + v reset, opt.py:513-520), ``Kmin > 0``
  (opt.py:578), an engaging ``gradg_clip`` (opt.py:539-542), ``rand_init`` (opt.py:432), ``kfac_rand=True``
  (opt.py:351-356);
* one epoch of the reference's own ``iter()`` (SGD with momentum + weight decay, Adam) against the fused step.
"""
import copy

import numpy as np
import pytest
import torch

from conftest import golden_vec_errors, load_golden, model_from_golden, rel_err, zoo_model_for_golden

pytestmark = pytest.mark.gpu

RTOL_VEC = 1e-4
RTOL_LAM = 1e-3


def _v_rand(P, seed=1226):
    gen = torch.Generator().manual_seed(seed + 3)
    vr = torch.randn(P, generator=gen, dtype=torch.float64)
    return vr / vr.norm()


@pytest.mark.parametrize("kind", ["chest_densenet121", "chest_vgg"])
def test_chest_models_match_reference_golden_at_config_batch(kind):
    """grad / Hv / vGHv of the full-size chest models at the config's batch of 4 against the reference's own fp32
    CPU result (2^17-index sample + classifier tail + norm per vector).  A mismatch above rtol is admissible only when
    the fp64 jet oracle, conditioned on the ReLU decisions the GPU took, reproduces the GPU vectors at rtol and every
    differing decision sits on a pre-activation that is undecidable in fp32 (tests/kinks.py)."""
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    g = load_golden(kind)
    model, loss, x, y = zoo_model_for_golden(kind, g)
    op = B200HVPOperator(model, [x, y], loss)
    P = sum(p.numel() for p in model.parameters())
    v0 = torch.from_numpy(np.ones(P) / np.sqrt(P))
    vr = _v_rand(P)
    from conftest import checksum
    assert np.allclose(checksum(vr.numpy()), g["v_rand_check"], rtol=1e-12)
    hv0 = op.Hv(v0, storedGrad=True).cpu().numpy()
    grad = op.stored_grad.cpu().numpy()
    hvr = op.Hv(vr.numpy(), storedGrad=True).cpu().numpy()
    vgr = op.vGHv(vr, storedGrad=True).cpu().numpy()
    checks = [("grad", "grad", None, grad), ("hv_v0", "hv", v0, hv0), ("hv_vrand", "hv", vr, hvr),
              ("vghv_vrand", "vghv", vr, vgr)]
    report = {name: golden_vec_errors(g, name, vec) for name, _, _, vec in checks}
    print(kind, report)
    # downstream of every kink the gradient is continuous in the decisions: the classifier tail must agree as is
    assert report["grad"]["tail"] < RTOL_VEC, report
    assert abs(float(op.loss_value) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    bad = [n for n, r in report.items() if max(r["sample"], r["norm"]) >= RTOL_VEC]
    if bad:
        from kinks import explain_by_kinks
        flips = explain_by_kinks(op, model, x, y, checks, rtol=RTOL_VEC)
        assert flips >= 1, "golden mismatch in %s without any ambiguous ReLU decision: %r" % (bad, report)
    # BatchNorm running statistics after one train-mode forward
    pnames = {n for n, _ in model.named_parameters()}
    buf = np.concatenate([t.detach().reshape(-1).double().cpu().numpy() for k, t in model.state_dict().items()
                          if k not in pnames])
    assert rel_err(buf, g["buffers_after_one_pass"]) < 1e-5


@pytest.mark.parametrize("kind", ["chest_densenet121", "chest_vgg"])
def test_chest_comp_rho_with_config_flags(kind, tmp_path):
    """comp_rho + comp_gradrho with rand_init=True, gradg_clip=100, eps 0.1 (chestxray_best_reg.py:120-134)."""
    from optwboundeigenval_b200.spectral import SpectralState
    g = load_golden(kind)
    meta = eval(str(g["meta"]))
    model, loss, x, y = zoo_model_for_golden(kind, g)
    st = SpectralState(model, loss, mu=0.01, K=0, pow_iter_eps=meta["eps"], max_pow_iter=meta["max_pow_iter"],
                       ignore_bad_vals=False, rand_init=True, gradg_clip=meta["clip"], verbose=True,
                       verbose_log_file=str(tmp_path / "v.log"), log_file=str(tmp_path / "l.log"))
    st.v = torch.zeros_like(st.v)                  # rand_init must ignore whatever self.v holds (opt.py:432)
    i, rn, size = st.comp_rho({"image": x, "label": y})
    traj_ref = g["rho1_traj"]
    rows = [ln.split("\t") for ln in open(tmp_path / "v.log").read().splitlines() if ln and ln[0].isdigit()]
    lam = np.array([float(r[1]) for r in rows])
    m = min(len(lam), len(traj_ref))
    print(kind, "iters", i, int(g["rho1_iters"]), "rho", st.rho, float(g["rho1_rho"]), "lam", lam[:m], traj_ref[:m, 1])
    assert size == 4
    assert abs(i - int(g["rho1_iters"])) <= 1
    # per-iteration lambda = v.Hv: the fp32-ambiguous ReLU / max-pool decisions of the full-size models move the whole Hv
    # vector by several 1e-3 (explained decision by decision in test_chest_models_match_reference_golden_at_config_batch),
    # and lambda inherits at most that relative error -- measured here on Hv at the start vector, factor 2 of slack
    P = st.ndim
    v0 = torch.from_numpy(np.ones(P) / np.sqrt(P))
    hv_err = golden_vec_errors(g, "hv_v0", st.hvp_op.Hv(v0, storedGrad=True).cpu().numpy())["sample"]
    lam_tol = max(5e-3, 2.0 * hv_err)
    print(kind, "hv_v0 sample error", hv_err, "-> lambda tolerance", lam_tol)
    # (the verbose log prints %f: 1e-6 absolute resolution; the first lambda, at the start vector, is ~2e-4)
    np.testing.assert_allclose(lam[:m], traj_ref[:m, 1], rtol=lam_tol, atol=5e-6)
    if i == int(g["rho1_iters"]):
        assert abs(st.rho - float(g["rho1_rho"])) <= lam_tol * float(g["rho1_rho"])
        st.g = max(0.0, st.rho - st.K, st.Kmin - st.rho)
        st.comp_gradrho()
        gn = float(torch.norm(st.gradrho))
        assert abs(gn - float(g["rho1_gradrho_norm"])) < 1e-6 * gn           # both clipped to 100
        err = golden_vec_errors(g, "rho1_gradrho", st.gradrho.cpu().numpy())
        print(kind, "gradrho", err)
        assert err["sample"] < 5e-2                                            # direction: v itself carries the noise


def _usps_state(g, **kw):
    from optwboundeigenval_b200.spectral import SpectralState
    meta = eval(str(g["meta"]))
    model, loss = model_from_golden("usps", g)
    st = SpectralState(model, loss, mu=0.01, K=meta.get("K", 0), Kmin=meta.get("Kmin", 0), pow_iter_eps=meta["eps"],
                       max_pow_iter=meta["max_pow_iter"], ignore_bad_vals=meta.get("ignore_bad_vals", False),
                       rand_init=meta.get("rand_init", False), gradg_clip=meta.get("clip"), **kw)
    return st, model, loss, meta


def test_ignore_bad_vals_kmin_clip_rand_init_match_reference():
    """usps_flags: 6 iterations at eps 1e-7 never converge -> rho = -1, v reset to ones/sqrt(P) (opt.py:513-520);
    Kmin = 0.5 -> g = Kmin - rho = 1.5 (opt.py:578); the penalty gradient at the reset vector, clipped to 1e-3."""
    g = load_golden("usps_flags")
    st, model, loss, meta = _usps_state(g)
    data = [torch.from_numpy(g["x"]), torch.from_numpy(g["y"])]
    i, rn, size = st.comp_rho(data)
    assert i == int(g["rho1_iters"]) == 5
    assert st.rho == -1 and float(g["rho1_rho"]) == -1
    assert rel_err(st.v.cpu().numpy(), g["rho1_v"]) < 1e-12                  # the reset vector
    assert abs(st.norm - float(g["rho1_norm"])) <= 1e-3 * float(g["rho1_norm"])
    st.g = np.max([0.0, st.rho - st.K, st.Kmin - st.rho])
    assert st.g == float(g["rho1_g"]) == 1.5
    st.comp_gradrho()
    assert abs(float(torch.norm(st.gradrho)) - float(g["rho1_gradrho_norm"])) < 1e-9      # clipped to 1e-3
    assert rel_err(st.gradrho.cpu().numpy(), g["rho1_gradrho"]) < 2e-4
    # second minibatch: rand_init restarts from ones/sqrt(P), again no convergence
    i2, _, _ = st.comp_rho([torch.from_numpy(g["x2"]), torch.from_numpy(g["y2"])])
    assert i2 == int(g["rho2_iters"]) and st.rho == float(g["rho2_rho"]) == -1


def test_kmin_branch_sign_and_fused_step_direction():
    """usps_kmin: K = 1e9, Kmin = 5 > rho -> g = Kmin - rho, sign = -1 (opt.py:633): the step is grad f - mu * grad rho."""
    g = load_golden("usps_kmin")
    st, model, loss, meta = _usps_state(g)
    data = [torch.from_numpy(g["x"]), torch.from_numpy(g["y"])]
    st.comp_g(data)
    assert abs(st.rho - float(g["rho1_rho"])) <= RTOL_LAM * float(g["rho1_rho"])
    assert abs(st.g - float(g["rho1_g"])) <= 1e-3 * float(g["rho1_g"])
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    before = torch.cat([q.detach().reshape(-1).clone() for q in model.parameters()]).double()
    st.fused_step(optimizer=opt)
    after = torch.cat([q.detach().reshape(-1) for q in model.parameters()]).double()
    assert abs(float(torch.norm(st.gradrho)) - float(g["rho1_gradrho_norm"])) < 1e-9      # clip engaged: 1e-3
    want = torch.from_numpy(g["rho1_gradf"]).double().cuda() - 0.01 * torch.from_numpy(g["rho1_gradrho"]).double().cuda()
    got = (before - after) / 0.5
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 2e-4
    assert rel_err(st.gradg.cpu().numpy(), -g["rho1_gradrho"]) < 2e-3


def test_kfac_rand_targets_match_reference():
    """usps_lobpcg_rand: kfac_rand=True draws the Fisher targets from the model's own output with torch's global CPU
    generator (opt.py:351-356); seeded identically, the preconditioned iteration reproduces the reference."""
    from optwboundeigenval_b200.spectral import SpectralState
    g = load_golden("usps_lobpcg_rand")
    meta = eval(str(g["meta"]))
    model, loss = model_from_golden("usps", g)
    st = SpectralState(model, loss, pow_iter_eps=meta["eps"], max_pow_iter=meta["max_pow_iter"], ignore_bad_vals=False,
                       lobpcg=True, kfac_batch=1, kfac_rand=True, pow_iter_alpha=lambda k: np.exp(-4 * k))
    torch.manual_seed(1226 + 11)
    i, rn, size = st.comp_rho([torch.from_numpy(g["x"]), torch.from_numpy(g["y"])])
    assert i == int(g["rho1_iters"])
    assert abs(st.rho - float(g["rho1_rho"])) <= RTOL_LAM * float(g["rho1_rho"])
    v = st.v.cpu().numpy()
    assert min(rel_err(v, g["rho1_v"]), rel_err(-v, g["rho1_v"])) < 1e-3


@pytest.mark.parametrize("name,kind", [("usps_iter_sgd", "usps"), ("usps_iter_adam", "usps"), ("cifar_iter_sgd", "cifar_densenet")])
def test_one_epoch_of_iter_matches_the_reference(name, kind, tmp_path):
    """The reference's own iter() (opt.py:580-763) over a few minibatches -- comp_g, penalty gradient, clip, step
    assembly, SGD(momentum 0.9, weight decay) / Adam update, epoch loss in evaluation mode, end-of-epoch rho --
    against spectral.iter_epoch with the fused step and the forward-only evaluation pass."""
    import random
    from optwboundeigenval_b200.spectral import SpectralState
    g = load_golden(name)
    meta = eval(str(g["meta"]))
    model, loss = model_from_golden(kind, g)
    model.cuda()
    if meta["optimizer"] == "sgd":
        optim = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    else:
        optim = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    st = SpectralState(model, loss, mu=meta.get("mu", 0.01), K=meta.get("K", 0), pow_iter_eps=meta["eps"],
                       max_pow_iter=meta["max_pow_iter"], ignore_bad_vals=False, gradg_clip=meta.get("clip"), verbose=True,
                       verbose_log_file=str(tmp_path / "v.log"), log_file=str(tmp_path / "l.log"))
    st.optimizer = optim
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    st.dataloader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=meta["batch"])
    random.seed(1226)
    st.iter()
    flat = np.concatenate([t.detach().reshape(-1).double().cpu().numpy() for t in model.state_dict().values()])
    state0 = g["state0"].astype(np.float64)
    want = g["state_after_iter"].astype(np.float64)
    # the parameter UPDATE of the epoch (not the parameters: the update is 1e-2 of them) within 1e-3
    upd_err = np.linalg.norm((flat - state0) - (want - state0)) / np.linalg.norm(want - state0)
    print(name, "update rel err", upd_err, "f", st.f, float(g["f"]), "rho", st.rho, float(g["rho"]))
    assert upd_err < 2e-3
    # DenseNet3: the epoch loss is an evaluation-mode forward on running statistics that have seen two minibatches, and
    # the end-of-epoch rho stops at eps = 5e-2 -- both amplify the 3e-4 difference of the updates
    loose = kind == "cifar_densenet"
    assert abs(st.f - float(g["f"])) <= (1e-3 if loose else 1e-4) * abs(float(g["f"]))
    assert abs(st.rho - float(g["rho"])) <= (5e-2 if loose else 5e-3) * float(g["rho"])
    assert abs(st.h - float(g["h"])) <= (2e-3 if loose else 1e-3) * abs(float(g["h"]))

    # the per-minibatch verbose line "j rho norm |grad f| |grad g|" (opt.py:715-719) is the numeric line that follows
    # comp_rho's closing "Power Iter Time ..." line
    def per_batch(text):
        lines, out = text.splitlines(), []
        for k in range(1, len(lines)):
            f = lines[k].split("\t")
            if lines[k - 1].startswith("Power Iter") and len(f) == 5 and f[0].strip().isdigit():
                out.append([float(t) for t in f[1:]])
        return np.array(out)
    mine, ref = per_batch(open(tmp_path / "v.log").read()), per_batch(str(g["verbose_log"]))
    n = meta["n_batches"]
    assert len(mine) >= n and len(ref) >= n
    np.testing.assert_allclose(mine[:n, [0, 2, 3]], ref[:n, [0, 2, 3]], rtol=5e-3, atol=1e-6)     # rho, |grad f|, |grad g|


@pytest.mark.parametrize("which", ["sgd_nesterov", "sgd_plain", "adam"])
def test_fused_optimizer_update_equals_torch_optim(which):
    """b2s_step_fused against torch.optim on the same gradients, several steps, state included."""
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.spectral import SpectralState
    model, loss = zoo.build("usps")
    model.train().cuda()
    twin = copy.deepcopy(model)

    def make(m):
        if which == "sgd_nesterov":
            return torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3)
        if which == "sgd_plain":
            return torch.optim.SGD(m.parameters(), lr=0.05, weight_decay=1e-3)
        return torch.optim.Adam(m.parameters(), lr=1e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=1e-3)
    opt_a, opt_b = make(model), make(twin)
    st = SpectralState(model, loss, mu=0.3, K=0.0, pow_iter_eps=1e-3, max_pow_iter=40, ignore_bad_vals=False, gradg_clip=0.5)
    for k in range(4):
        x, y = zoo.synthetic_batch("usps", 32, seed=50 + k)
        st.comp_g([x, y])
        assert st.g > 0
        st.fused_step(optimizer=opt_a)
        # the same step direction through the reference's assembly (fp64 p, per-parameter fp32 slices) and torch's update
        p = st.gradf + 0.3 * st.gradg
        i = 0
        for q in twin.parameters():
            q.grad = p[i:i + q.numel()].view(q.shape).float()
            i += q.numel()
        opt_b.step()
        a = torch.cat([q.detach().reshape(-1) for q in model.parameters()])
        b = torch.cat([q.detach().reshape(-1) for q in twin.parameters()])
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-6, (which, k)
        for qa, qb in zip(model.parameters(), twin.parameters()):
            assert torch.allclose(qa.grad, qb.grad, rtol=1e-6, atol=1e-12)
        # the twin's parameters become this model's for the next minibatch (identical trajectories)
    sa, sb = opt_a.state_dict()["state"], opt_b.state_dict()["state"]
    assert sa.keys() == sb.keys()
    for k in sa:
        for key in sa[k]:
            va, vb = sa[k][key], sb[k][key]
            assert torch.allclose(va.float().cpu(), vb.float().cpu(), rtol=1e-5, atol=1e-9), (which, k, key)


def test_eval_pass_matches_torch_eval_forward():
    """comp_f (opt.py:544-572): loss and output in evaluation mode (BatchNorm running statistics) on a BatchNorm model
    and on a model with a softmax tail; running statistics untouched; the next comp_rho still works."""
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.spectral import SpectralState
    for kind, batch in (("cifar_densenet", 16), ("usps", 40), ("chest_densenet121", 2)):
        model, loss = zoo.build(kind)
        model.train()
        x, y = zoo.synthetic_batch(kind, batch)
        with torch.no_grad():                          # move the running statistics away from their initial values
            model(x)
        ref = copy.deepcopy(model).eval()
        with torch.no_grad():
            out_ref = ref(x)
            f_ref = float(loss(out_ref, y))
        st = SpectralState(model, loss, pow_iter_eps=1e-2, max_pow_iter=3, ignore_bad_vals=False)
        before = [b.clone() for b in model.buffers()]
        f, out = st.comp_f(x, y)
        assert abs(f - f_ref) <= 2e-5 * abs(f_ref), (kind, f, f_ref)
        assert rel_err(out.cpu().numpy(), out_ref.numpy()) < 2e-5, kind
        assert all(torch.equal(a, b) for a, b in zip(before, model.buffers()))
        st.comp_rho([x, y])                            # a base pass after the evaluation pass
        assert st.rho > 0


def test_operators_sharing_a_plan_stay_independent():
    """Two operators of one model share a plan; interleaving them must behave like the reference's independent objects."""
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    model, loss = zoo.build("cifar_densenet")
    model.train()
    P = sum(p.numel() for p in model.parameters())
    v = torch.from_numpy(np.ones(P) / np.sqrt(P))
    xa, ya = zoo.synthetic_batch("cifar_densenet", 8, seed=1)
    xb, yb = zoo.synthetic_batch("cifar_densenet", 8, seed=2)
    a = B200HVPOperator(model, [xa, ya], loss)
    hva = a.Hv(v, storedGrad=True).clone()
    b = B200HVPOperator(model, [xb, yb], loss)
    hvb = b.Hv(v, storedGrad=True).clone()
    assert rel_err(hva.cpu().numpy(), hvb.cpu().numpy()) > 1e-2
    buffers = [t.clone() for t in model.buffers()]
    again = a.Hv(v, storedGrad=True)                   # b's base pass is cached in the plan: a must get its own back
    assert rel_err(again.cpu().numpy(), hva.cpu().numpy()) < 1e-5
    assert all(torch.equal(s, t) for s, t in zip(buffers, model.buffers()))       # no extra running-stat update
    assert rel_err(b.vGHv(v, storedGrad=True).cpu().numpy(), b.vGHv(v, storedGrad=True).cpu().numpy()) < 1e-5


def test_hundred_class_head_and_invalid_label():
    """CIFAR-100 sized head (cifar100_ResNet_mu0.py: 100 classes) against the autograd oracle; a label outside [0, C)
    poisons the loss instead of reading out of bounds."""
    import torch.nn as nn
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    from oracle import autograd_oracle as ao
    torch.manual_seed(11)
    model = nn.Sequential(nn.Flatten(), nn.Linear(48, 64), nn.ReLU(), nn.Linear(64, 100)).train()
    loss = nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(24, 3, 4, 4, generator=g)
    y = torch.randint(0, 100, (24,), generator=g)
    P = sum(p.numel() for p in model.parameters())
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v /= v.norm()
    ref = ao.AutogradSpectralOperator(copy.deepcopy(model), [x, y], loss)
    op = B200HVPOperator(model, [x, y], loss)
    hv = op.Hv(v, storedGrad=True)
    assert rel_err(op.stored_grad.cpu().numpy(), ref.gradient().detach().numpy()) < RTOL_VEC
    assert rel_err(hv.cpu().numpy(), ref.hv(v).numpy()) < RTOL_VEC
    assert rel_err(op.vGHv(v, storedGrad=True).cpu().numpy(), ref.vghv(v).numpy()) < RTOL_VEC
    bad = y.clone()
    bad[3] = 100
    op2 = B200HVPOperator(model, [x, bad], loss)
    op2.Hv(v, storedGrad=True)
    assert torch.isnan(op2.loss_value).all()


def test_small_model_loop_runs_without_host_round_trips():
    """forest MLP: the whole eigen-iteration is one graph launch (WHILE node); same trajectory as the polled loop."""
    import os
    from optwboundeigenval_b200 import _lib, zoo
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator, clear_plans
    model, loss = zoo.build("forest")
    model.train()
    x, y = zoo.synthetic_batch("forest", 128)
    P = sum(p.numel() for p in model.parameters())
    v0 = torch.from_numpy(np.ones(P) / np.sqrt(P))
    res = {}
    for mode in ("1", "0"):
        os.environ["B2S_DEVICE_LOOP"] = mode
        clear_plans()
        op = B200HVPOperator(model, [x, y], loss)
        out = op.power_iterate(v0, 1e-3, 1000, want_trajectory=True)
        res[mode] = (out.iters, out.lam, out.trajectory.copy(), out.v.cpu().numpy())
        # a second call reuses the instantiated loop
        out2 = op.power_iterate(v0, 1e-3, 1000)
        assert out2.iters == out.iters
    os.environ.pop("B2S_DEVICE_LOOP")
    clear_plans()
    assert res["1"][0] == res["0"][0]
    # same arithmetic, different interleaving of the fp32 weight-gradient atomics: residual norms near 1e-5 move by 1e-9
    np.testing.assert_allclose(res["1"][2], res["0"][2], rtol=1e-3, atol=1e-7)
    assert rel_err(res["1"][3], res["0"][3]) < 1e-5
