"""CPU-only checks of the boundary: the library builds, loads, exports every symbol the header
declares, and the product path refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_spectral.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from optwboundeigenval_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    # the ctypes table binds exactly the declared functions
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().b2s_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_header():
    from optwboundeigenval_b200 import tracer, _lib
    assert ctypes.sizeof(tracer.CTensor) == 32
    assert ctypes.sizeof(tracer.COp) == 72
    assert ctypes.sizeof(_lib.PowerCfg) == 32
    assert ctypes.sizeof(_lib.PowerResult) == 64


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    from optwboundeigenval_b200.spectral import SpectralState
    model, loss = zoo.build("forest")
    x, y = zoo.synthetic_batch("forest", 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        B200HVPOperator(model, [x, y], loss)
    with pytest.raises(RuntimeError, match="CUDA"):
        SpectralState(model, loss)
    # the C ABI itself also refuses
    from optwboundeigenval_b200 import _lib, tracer
    lib = _lib.load()
    tape = tracer.trace(model, loss, (54,))
    h = ctypes.c_void_p()
    t, o = tape.c_tensors(), tape.c_ops()
    b = (ctypes.c_int64 * len(tape.buf_elems))(*tape.buf_elems)
    rc = lib.b2s_plan_create(t, len(t), b, len(b), o, len(o), tape.logits, tape.head, tape.n_params, 8, 0,
                             ctypes.byref(h))
    assert rc != 0 and b"CUDA" in lib.b2s_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "optwboundeigenval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_unsupported_modules_fail_loudly():
    from optwboundeigenval_b200 import tracer
    import torch.nn as nn
    m = nn.Sequential(nn.Linear(4, 4), nn.Tanh(), nn.Linear(4, 2))
    with pytest.raises(tracer.UnsupportedModel, match="Tanh"):
        tracer.trace(m, nn.CrossEntropyLoss(), (4,))
    with pytest.raises(tracer.UnsupportedModel, match="MSELoss"):
        tracer.trace(nn.Sequential(nn.Linear(4, 2)), nn.MSELoss(), (4,))
