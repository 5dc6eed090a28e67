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
