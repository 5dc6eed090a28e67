"""ctypes binding of libb200spectral.so (include/b200_spectral.h).

There is deliberately no fallback: if the shared library is missing, cannot be
loaded, or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int32, c_int64, c_void_p

from .tracer import COp, CTensor

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libb200spectral.so")
ABI_VERSION = 1


class PowerCfg(ctypes.Structure):
    _fields_ = [("max_iter", c_int32), ("eps", c_double), ("h_alpha", POINTER(c_double)), ("precond", c_int32)]


class PowerResult(ctypes.Structure):
    _fields_ = [("iters", c_int32), ("converged", c_int32), ("lam", c_double), ("norm", c_double),
                ("rn", c_double), ("vnn", c_double), ("stop", c_double * 3)]


class ProfEntry(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("launches", c_int32), ("ms", c_double), ("flops", c_double),
                ("bytes", c_double)]


class StepOpt(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("first_step", c_int32), ("nesterov", c_int32), ("maximize", c_int32),
                ("write_gradrho", c_int32), ("lr", c_double), ("momentum", c_double), ("dampening", c_double),
                ("weight_decay", c_double), ("beta1", c_double), ("beta2", c_double), ("eps", c_double),
                ("step_size", c_double), ("bias2_sqrt", c_double)]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "b2s_abi_version": (c_int32, []),
    "b2s_last_error": (c_char_p, []),
    "b2s_launch_count": (c_int64, []),
    "b2s_set_tensor_core_mode": (c_int32, [c_int32]),
    "b2s_plan_create": (c_int32, [POINTER(CTensor), c_int32, POINTER(c_int64), c_int32, POINTER(COp), c_int32,
                                  c_int32, c_int32, c_int64, c_int32, c_int32, POINTER(c_void_p)]),
    "b2s_plan_destroy": (c_int32, [c_void_p]),
    "b2s_plan_set_stream": (c_int32, [c_void_p, c_void_p]),
    "b2s_plan_set_graphs": (c_int32, [c_void_p, c_int32]),
    "b2s_plan_workspace_bytes": (c_int64, [c_void_p]),
    "b2s_plan_set_bn_third_order": (c_int32, [c_void_p, c_int32]),
    "b2s_plan_set_bn_buffers": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p]),
    "b2s_base_pass": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_double, c_void_p,
                                c_void_p]),
    "b2s_hv": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "b2s_eval_pass": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_double, c_void_p,
                                c_void_p]),
    "b2s_vghv": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "b2s_debug_read": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "b2s_profile_pass": (c_int32, [c_void_p, c_int32, c_int32, POINTER(ProfEntry), c_int32, POINTER(c_int32)]),
    "b2s_grad_f32": (c_void_p, [c_void_p]),
    "b2s_hv_f32": (c_void_p, [c_void_p]),
    "b2s_power_iterate": (c_int32, [c_void_p, c_void_p, POINTER(PowerCfg), POINTER(PowerResult), c_void_p]),
    "b2s_pi_create": (c_int32, [c_int64, c_int32, c_int32, POINTER(c_void_p)]),
    "b2s_pi_destroy": (c_int32, [c_void_p]),
    "b2s_pi_reset": (c_int32, [c_void_p, c_void_p, POINTER(PowerCfg), c_void_p]),
    "b2s_pi_v32": (c_void_p, [c_void_p]),
    "b2s_pi_step": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "b2s_pi_residual": (c_void_p, [c_void_p, c_void_p]),
    "b2s_pi_precond_update": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "b2s_pi_done": (c_int32, [c_void_p, c_void_p]),
    "b2s_pi_result": (c_int32, [c_void_p, POINTER(PowerResult), c_void_p, c_void_p, c_void_p]),
    "b2s_kfac_dims": (c_int32, [c_void_p, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "b2s_kfac_build": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "b2s_kfac_clear": (c_int32, [c_void_p]),
    "b2s_kfac_set": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p]),
    "b2s_kfac_apply": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "b2s_step_assemble": (c_int32, [c_void_p, c_void_p, ctypes.c_double, c_int64, c_void_p, c_void_p, c_void_p]),
    "b2s_clip_norm": (c_int32, [c_void_p, c_int64, c_double, c_void_p, c_void_p, c_void_p]),
    "b2s_clip_scratch_doubles": (c_int32, []),
    "b2s_step_fused": (c_int32, [c_void_p, c_void_p, c_double, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, POINTER(StepOpt), c_void_p]),
    "b2s_comm_unique_id": (c_int32, [c_void_p]),
    "b2s_comm_init": (c_int32, [c_void_p, c_void_p, c_int32, c_int32]),
    "b2s_comm_destroy": (c_int32, [c_void_p]),
    "b2s_plan_set_global_batch": (c_int32, [c_void_p, c_int64]),
    "b2s_comm_peer_local": (c_int32, [c_void_p, c_void_p]),
    "b2s_comm_peer_attach": (c_int32, [c_void_p, c_void_p]),
    "b2s_comm_peer_ready": (c_int32, [c_void_p]),
    "b2s_comm_peer_disable": (c_int32, [c_void_p]),
}

_lib = None


def load():
    """Load the library or raise -- never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libb200spectral.so is not built (%s). Run `python -m optwboundeigenval_b200.build`; "
            "this package has no CPU or autograd fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.b2s_abi_version() != ABI_VERSION:
        raise RuntimeError("libb200spectral ABI %d != expected %d" % (lib.b2s_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int, what: str = "libb200spectral"):
    if rc != 0:
        msg = load().b2s_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(load().b2s_launch_count())
