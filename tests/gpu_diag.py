"""One-shot GPU diagnostic (run under gpurun): prints parity numbers for every config and, when a
pass disagrees with the CPU oracles, localises the first op whose cached tensor differs from the
jet oracle.  Test infrastructure (imports oracle/)."""
import copy
import ctypes
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from optwboundeigenval_b200 import _lib, tracer, zoo                      # noqa: E402
from optwboundeigenval_b200.hvp_operator import B200HVPOperator, SpectralPlan, flat_parameters   # noqa: E402
from optwboundeigenval_b200.spectral import SpectralState                # noqa: E402
from oracle import autograd_oracle as ao                                   # noqa: E402
from oracle.jet_oracle import JetTapeOracle                                # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def read_tensor(plan, adjoint, order, t, batch):
    vt = plan.tape.tensors[t]
    out = np.zeros((batch,) + vt.shape, dtype=np.float32)
    _lib.check(plan.lib.b2s_debug_read(plan.handle, adjoint, order, t, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def count_mask_flips(plan, cpu_model, x, batch):
    """ReLU on/off decisions that differ between the GPU forward and torch's fp32 CPU forward."""
    tape = plan.tape
    acts = {}
    hooks = []
    for name, m in cpu_model.named_modules():
        if isinstance(m, torch.nn.ReLU):
            hooks.append(m.register_forward_hook(lambda mod, i, o, name=name: acts.setdefault(name, []).append(o.detach().clone())))
    with torch.no_grad():
        cpu_model(x)
    for h in hooks:
        h.remove()
    cpu_list = [a for name in acts for a in acts[name]]
    gpu_ops = [op for op in tape.ops if op.flags & tracer.F_RELU]
    total = flips = 0
    k = 0
    for op in gpu_ops:
        got = read_tensor(plan, 0, 0, op.out, batch)
        cands = [a for a in cpu_list if tuple(a.shape[1:]) == got.shape[1:]]
        if k >= len(cpu_list):
            break
        want = cpu_list[k].numpy()
        k += 1
        if want.shape != got.shape:
            continue
        d = (got > 0) != (want > 0)
        total += d.size
        flips += int(d.sum())
    print("   ReLU decisions compared: %d, differing: %d" % (total, flips))


def localise(plan, jo, batch, orders=(0, 1, 2), tol=1e-3):
    tape = plan.tape
    for K in orders:
        for oi, op in enumerate(tape.ops):
            got = read_tensor(plan, 0, K, op.out, batch)
            want = jo.view(jo.fw, K, op.out).numpy()
            e = rel(got, want) if np.linalg.norm(want) > 0 else float(np.linalg.norm(got))
            if e > tol:
                print("   first forward mismatch: order %d op %d %s (%s) rel=%.3e" % (K, oi, op.name, op.kind, e))
                break
        for oi in range(len(tape.ops) - 1, -1, -1):
            op = tape.ops[oi]
            if op.flags & tracer.F_FIRST:
                continue
            got = read_tensor(plan, 1, K, op.inp, batch)
            want = jo.view(jo.bw, K, op.inp).numpy()
            e = rel(got, want) if np.linalg.norm(want) > 0 else float(np.linalg.norm(got))
            if e > tol:
                print("   first backward mismatch (input adjoint): order %d op %d %s (%s) rel=%.3e" % (K, oi, op.name, op.kind, e))
                break


def run_config(kind, batch, graphs=True, do_vghv=True, localise_always=False):
    print("=== %s batch=%d graphs=%s" % (kind, batch, graphs), flush=True)
    model, loss = zoo.build(kind)
    model.train()
    if kind != "forest" and kind != "usps" and os.environ.get("DIAG_RANDOM_BN", "0") == "1":
        g = torch.Generator().manual_seed(5)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.data = 1 + 0.3 * torch.randn(m.weight.shape, generator=g)
                m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
    x, y = zoo.synthetic_batch(kind, batch)
    cpu_model = copy.deepcopy(model)
    t0 = time.time()
    ref = ao.AutogradSpectralOperator(cpu_model, [x, y], loss)
    P = sum(p.numel() for p in model.parameters())
    gen = torch.Generator().manual_seed(11)
    v = torch.randn(P, generator=gen, dtype=torch.float64)
    v /= v.norm()
    g_ref = ref.gradient().detach()
    hv_ref = ref.hv(v)
    vg_ref = ref.vghv(v) if do_vghv else None
    print("   cpu oracle: %.1fs" % (time.time() - t0), flush=True)
    if os.environ.get("DIAG_FP64", "0") == "1":
        m64 = copy.deepcopy(cpu_model).double()
        r64 = ao.AutogradSpectralOperator(m64, [x.double(), y], loss)
        g64 = r64.gradient().detach(); h64 = r64.hv(v.float().double())
        print("   fp32 autograd vs fp64 autograd: grad %.3e hv %.3e" % (rel(g_ref, g64), rel(hv_ref, h64)))
        g_ref, hv_ref = g64, h64
        print("   (GPU numbers below are against the fp64 truth)")

    op = B200HVPOperator(model, [x, y], loss)
    hv = op.Hv(v, storedGrad=True)
    if not graphs:
        _lib.check(op.plan.lib.b2s_plan_set_graphs(op.plan.handle, 0))
    torch.cuda.synchronize()
    print("   grad  rel=%.3e   loss gpu=%.8f cpu=%.8f" % (rel(op.stored_grad.cpu(), g_ref), float(op.loss_value), ref.loss_value))
    print("   hv    rel=%.3e" % rel(hv.cpu(), hv_ref))
    if rel(op.stored_grad.cpu(), g_ref) > 1e-5:
        off = 0
        rows = []
        ga = op.stored_grad.cpu().numpy(); gb = g_ref.numpy()
        ha = hv.cpu().numpy(); hb = hv_ref.numpy()
        for name, prm in model.named_parameters():
            k = prm.numel()
            rows.append((rel(ga[off:off + k], gb[off:off + k]), rel(ha[off:off + k], hb[off:off + k]), name, tuple(prm.shape),
                         float(np.linalg.norm(gb[off:off + k]))))
            off += k
        print("   per-parameter grad/hv errors (in order):")
        for e, eh, name, shp, nrm in rows:
            print("      %-40s %-18s grad %.2e  hv %.2e  |g|=%.2e" % (name, shp, e, eh, nrm))
    if os.environ.get("DIAG_FLIPS", "0") == "1":
        cm = copy.deepcopy(cpu_model)
        count_mask_flips(op.plan, cm, x, batch)
    hv2 = op.Hv(v.numpy(), storedGrad=True)
    print("   hv(2) rel=%.3e (replay)" % rel(hv2.cpu(), hv_ref))
    bad = rel(op.stored_grad.cpu(), g_ref) > 1e-4 or rel(hv.cpu(), hv_ref) > 1e-4
    if do_vghv:
        vg = op.vGHv(v, storedGrad=True)
        torch.cuda.synchronize()
        print("   vghv  rel=%.3e" % rel(vg.cpu(), vg_ref))
        bad = bad or rel(vg.cpu(), vg_ref) > 1e-4
    # BN running stats side effect
    if any(isinstance(m, torch.nn.BatchNorm2d) for m in model.modules()):
        a = torch.cat([m.running_var.flatten() for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]).cpu()
        b = torch.cat([m.running_var.flatten() for m in cpu_model.modules() if isinstance(m, torch.nn.BatchNorm2d)])
        print("   running_var rel=%.3e" % rel(a, b))
    if bad or localise_always:
        tape = op.plan.tape
        head = tape.head
        if head in (tracer.HEAD_WBCE, tracer.HEAD_SIGMOID_WBCE):
            t, coef = SpectralPlan.wbce_coefficients(y.float())
            jo = JetTapeOracle(tape, ao.flat_params(cpu_model), x, t, coef, loss_scale=1.0)
        else:
            jo = JetTapeOracle(tape, ao.flat_params(cpu_model), x, y)
        jo.run(0); jo.run(1, v)
        if do_vghv:
            jo.run(2)
            # the library ran: hv(v) -> pass 2 (-> correction overwrites bw[2])
        print("   jet-oracle vs autograd: hv rel=%.3e" % rel(jo.out[1], hv_ref))
        # re-run passes on the GPU so caches correspond (vghv's correction sweep clobbers bw[2])
        op.Hv(v, storedGrad=True)
        localise(op.plan, jo, batch, orders=(0, 1))
    return op


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    only = sys.argv[1:] or ["forest", "usps", "cifar_densenet"]
    for kind in only:
        try:
            b = {"forest": 128, "usps": 64, "cifar_densenet": 8, "chest_vgg": 2, "chest_densenet121": 2}[kind]
            run_config(kind, b)
        except Exception:   # noqa: BLE001
            traceback.print_exc()
    # power iteration on forest
    try:
        model, loss = zoo.build("forest")
        x, y = zoo.synthetic_batch("forest", 128)
        st = SpectralState(model, loss, pow_iter_eps=1e-3, max_pow_iter=1000, ignore_bad_vals=False)
        t0 = time.time()
        i, rn, size = st.comp_rho([x, y])
        torch.cuda.synchronize()
        print("forest comp_rho: iters=%d rho=%.8g norm=%.4g rn=%.4g  %.3fs" % (i, st.rho, st.norm, rn, time.time() - t0))
        import copy
        cpu = copy.deepcopy(model).cpu()
        r = ao.power_iteration(ao.AutogradSpectralOperator(cpu, [x, y], loss).hv, ao.start_vector(st.ndim), eps=1e-3, max_iter=1000)
        print("forest oracle  : iters=%d rho=%.8g norm=%.4g" % (r["iters"], r["rho"], r["norm"]))
        print("launches so far:", _lib.launch_count())
    except Exception:   # noqa: BLE001
        traceback.print_exc()


if __name__ == "__main__":
    main()
