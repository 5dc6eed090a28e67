"""HBM roofline of the two fused vector kernels of the eigen-iteration (pi_dot + pi_update) at
P = 19 459 150 (VGG16-bn) and P = 2^26, through the C ABI.  52 bytes per element per iteration."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optwboundeigenval_b200 import _lib                                   # noqa: E402


def measure(n, iters=50, warmup=5):
    lib = _lib.load()
    st = ctypes.c_void_p()
    _lib.check(lib.b2s_pi_create(n, iters + warmup + 8, 0, ctypes.byref(st)))
    try:
        v0 = torch.full((n,), 1.0 / n ** 0.5, dtype=torch.float64, device="cuda")
        hv = torch.randn(n, dtype=torch.float32, device="cuda")
        cfg = _lib.PowerCfg()
        cfg.max_iter, cfg.eps, cfg.precond = iters + warmup + 4, 0.0, 0
        _lib.check(lib.b2s_pi_reset(st, ctypes.c_void_p(v0.data_ptr()), ctypes.byref(cfg), None))
        for _ in range(warmup):
            _lib.check(lib.b2s_pi_step(st, ctypes.c_void_p(hv.data_ptr()), None))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            _lib.check(lib.b2s_pi_step(st, ctypes.c_void_p(hv.data_ptr()), None))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    finally:
        lib.b2s_pi_destroy(st)
    bytes_ = 52.0 * n
    return {"n": n, "ms_per_iteration": ms, "bytes_per_iteration": bytes_, "GBps": bytes_ / ms / 1e6}


if __name__ == "__main__":
    peak = 6549.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    for n in (19459150, 1 << 26):
        r = measure(n)
        r["frac_of_measured_hbm_peak"] = r["GBps"] / peak
        print(json.dumps(r))
