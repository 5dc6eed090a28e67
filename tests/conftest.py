import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture %s missing" % path)
    return np.load(path, allow_pickle=False)


def model_from_golden(kind, gold, state_key="state0"):
    """zoo model carrying exactly the weights/buffers the reference model had."""
    import torch
    from optwboundeigenval_b200 import zoo
    model, loss = zoo.build(kind)
    flat = gold[state_key]
    sd = model.state_dict()
    names = [str(s) for s in gold["state_names"]]
    assert names == list(sd.keys()), "zoo model state_dict keys differ from the reference's"
    j = 0
    new = {}
    for k, v in sd.items():
        n = v.numel()
        new[k] = torch.from_numpy(np.asarray(flat[j:j + n])).to(v.dtype).view(v.shape)
        j += n
    model.load_state_dict(new)
    model.train()
    return model, loss


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ---- compact fixtures of the chest models (oracle/make_golden.py::_Packer) ---------------------------------
SAMPLE = 1 << 17


def sample_index(n):
    """the fixed index set the compact fixtures were sampled at (restates oracle/make_golden.py::sample_index)"""
    if n <= SAMPLE:
        return np.arange(n)
    return np.sort(np.random.RandomState(20261018).choice(n, SAMPLE, replace=False))


def checksum(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    w = np.cos(np.arange(a.size, dtype=np.float64) * 0.61803398875)
    return np.array([a.sum(), np.square(a).sum(), np.dot(a, w)])


def golden_vec_errors(gold, key, vec):
    """relative errors of a full vector against a golden entry: {'full'} for plain fixtures, {'sample', 'tail', 'norm'}
    for compact ones"""
    vec = np.asarray(vec, dtype=np.float64).reshape(-1)
    if key in gold:
        return {"full": rel_err(vec, gold[key])}
    idx = sample_index(vec.size)
    tail = gold[key + "_tail"]
    return {"sample": rel_err(vec[idx], gold[key + "_s"]), "tail": rel_err(vec[-tail.size:], tail),
            "norm": abs(np.linalg.norm(vec) - float(gold[key + "_norm"])) / float(gold[key + "_norm"])}


def zoo_model_for_golden(kind, gold):
    """zoo model + batch for a fixture; compact fixtures carry check sums instead of weights / inputs: the seeded zoo
    build must reproduce the reference's weights bit for bit (make_golden asserted it when the fixture was made)"""
    import torch
    from optwboundeigenval_b200 import zoo
    if "state0" in gold:
        model, loss = model_from_golden(kind, gold)
        return model, loss, torch.from_numpy(gold["x"]), torch.from_numpy(gold["y"])
    model, loss = zoo.build(kind)
    model.train()
    flat = np.concatenate([t.detach().reshape(-1).double().numpy() for t in model.state_dict().values()]).astype(np.float32)
    assert np.allclose(checksum(flat), gold["state0_check"], rtol=1e-12, atol=0), "zoo weights differ from the fixture's"
    batch = gold["y"].shape[0]
    x, y = zoo.synthetic_batch(kind, batch)
    assert np.allclose(checksum(x.numpy()), gold["x_check"], rtol=1e-12, atol=0)
    assert np.array_equal(y.numpy(), gold["y"])
    return model, loss, x, y
