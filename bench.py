#!/usr/bin/env python
"""bench.py -- HVPs/sec of the spectral-radius hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): params/cifar10_DenseNet_mu0_01_K10.py -- DenseNet3(40,12) on synthetic
32x32 images, 32 images per GPU (the reference's batch size), power iteration for lambda_max.
A *step* is one iteration of comp_rho's loop (opt.py:447-498) in steady state: one Hessian-vector
product of the cached minibatch (HVPOperator.Hv(v, storedGrad=True), opt.py:77-108) plus the
vector algebra and stopping test; the base pass (prepare_grad) is cached exactly as in the reference.

value   device-resident throughput: b2s_power_iterate run for exactly K iterations (eps = 0).
        With N GPUs each rank holds its own 32-image shard (weak scaling, global batch 32 N, synced
        BatchNorm statistics, one all-reduce of the P-vector per HVP); value counts 32-image-minibatch
        HVP equivalents: N * K / time (config.unit_note).
e2e     the same HVP through the reference-facing call B200HVPOperator.Hv(vec) with a HOST vector:
        per step 8P bytes host->device (pinned) and 8P bytes device->host are inside the timed region.
roofline / cpu_baseline / parity / per_config / gpu_autograd_yardstick: see DESIGN.md "Measurement".

--impl reference times the reference's own CPU implementation on the same config, rank 0 only: the
UNMODIFIED reference (opt.OptWBoundEignVal.comp_rho driving opt.HVPOperator, nested torch.autograd on
all host cores) when a checkout is present (baseline/_ref/optWBoundEigenval travels to the GPU box),
else the oracle port.  Warm-up = the first W iterations of comp_rho's loop (they include the base
pass), timed = the next K iterations, each with the loop's own residual / stopping-test algebra.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "HVPs/sec (power-iter lambda_max)"
UNIT = "HVP/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons while the timed region runs: NVML (about 1 ms per sample) when the
    bindings are importable, else nvidia-smi (about 100 ms per sample)"""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]          # nvmlClocksEventReason*: HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.mhz = []
        self.max_mhz = None
        self.seen = set()
        self.stop_flag = False
        self.source = "nvidia-smi"

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        self.source = "nvml"
        while not self.stop_flag:
            self.mhz.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            for n, b in zip(self.NAMES, self.BITS):
                if mask & b:
                    self.seen.add(n)
            time.sleep(0.002)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    if parts[0].replace(".", "").isdigit():
                        self.mhz.append(float(parts[0]))
                    if parts[1].replace(".", "").isdigit():
                        self.max_mhz = float(parts[1])
                    for i, n in enumerate(self.NAMES):
                        if parts[2 + i].lower().startswith("active"):
                            self.seen.add(n)
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:   # noqa: BLE001
            self._run_smi()

    def summary(self):
        if not self.mhz:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        import statistics
        return {"sm_mhz": statistics.median(self.mhz), "sm_max_mhz": self.max_mhz,
                "reasons": [n for n in self.NAMES if n in self.seen], "samples": len(self.mhz), "source": self.source}


def _ncu_summary():
    """dram bytes per launch of the dominant kernels from the committed ncu --set full captures (profiles/)"""
    for name in ("r2_ncu_top_kernel.json", "r1_ncu_top_kernel.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as fh:
                return json.load(fh), "profiles/" + name
    return None, None


def vector_roofline(peaks):
    """HBM roofline of the eigen-iteration's two fused vector kernels at P = 2^26 (DenseNet3's P = 176 122
    sits in L2, so the >= 90 % claim is measured at a DRAM-resident size; 52 algorithmic bytes per element)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_vec
    r = bench_vec.measure(1 << 26, iters=30, warmup=5)
    return {"kernel": "pi_dot_kernel + pi_update_kernel", "P": r["n"], "bound": "hbm", "achieved": r["GBps"],
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": r["GBps"] / peaks["hbm_gbs"],
            "ms_per_iteration": r["ms_per_iteration"], "algorithmic_bytes_per_iteration": r["bytes_per_iteration"]}


def measure_tf32_peak():
    """dense TF32 matmul rate of this GPU the way MEASURED_PEAKS.json measured bf16: torch.matmul 8192^3 with
    allow_tf32, best of 10 (library call, used only as the roofline denominator)"""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ---------------------------------------------------------------------------------------------------------
# the reference's own implementation (CPU, or the same stock autograd path on the GPU as a yard-stick)
# ---------------------------------------------------------------------------------------------------------
def _reference_objects(kind, like_model, use_gpu=False):
    """(opt module, reference model carrying like_model's weights, reference loss) or None without a checkout"""
    from oracle import reference_access as ra
    if ra.find_reference() is None:
        return None
    opt = ra.import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        model, loss = ra.ref_model_with_state(kind, like_model)
    model.train()
    return opt, model, loss


def reference_comp_rho_rate(kind, batch, warmup, steps, use_gpu=False):
    """HVPs/sec inside the UNMODIFIED reference's comp_rho (opt.py:418-533): warm-up = its first `warmup`
    iterations (incl. the base pass), timed = the next `steps` iterations.  Returns (rate, seconds, kind_str)."""
    import torch
    from optwboundeigenval_b200 import zoo
    zmodel, zloss = zoo.build(kind)
    x, y = zoo.synthetic_batch(kind, batch)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = _reference_objects(kind, zmodel, use_gpu)
    n_it = warmup + steps + 1
    if ref is None:                                   # no checkout on this machine: the oracle port
        from oracle import autograd_oracle as ao
        zmodel.train()
        op = ao.AutogradSpectralOperator(zmodel, [x, y], zloss)
        stamps = []

        def hv(v):
            stamps.append(time.time())
            return op.hv(v)
        P = sum(p.numel() for p in zmodel.parameters())
        ao.power_iteration(hv, ao.start_vector(P), eps=0.0, max_iter=n_it)
        dt = stamps[warmup + steps] - stamps[warmup]
        return steps / dt, dt, "port"
    opt, model, loss = ref
    stamps = []

    class Stamped(opt.HVPOperator):                   # time stamps only; the arithmetic is the reference's
        def Hv(self, vec, storedGrad=False):
            if use_gpu:
                torch.cuda.synchronize()
            stamps.append(time.time())
            return super().Hv(vec, storedGrad)

    real = opt.HVPOperator
    opt.HVPOperator = Stamped
    cwd = os.getcwd()
    import tempfile
    tmp = tempfile.mkdtemp(prefix="b2s_ref_")
    os.makedirs(os.path.join(tmp, "logs"))
    os.chdir(tmp)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            o = opt.OptWBoundEignVal(model, loss, torch.optim.SGD(model.parameters(), lr=0.1), mu=0.01, K=10,
                                     pow_iter_eps=0.0, max_pow_iter=n_it, use_gpu=use_gpu, ignore_bad_vals=False,
                                     header="bench", batch_size=batch)
            o.comp_rho([x, y])
    finally:
        os.chdir(cwd)
        opt.HVPOperator = real
    dt = stamps[warmup + steps] - stamps[warmup]
    return steps / dt, dt, "reference"


def reference_gpu_hv_rate(kind, batch, warmup, steps):
    """HVPs/sec of the UNMODIFIED opt.HVPOperator with use_gpu=True on this GPU (stock torch.autograd double backward
    through cuBLAS / cuDNN), driven by a plain normalised power iteration.  comp_rho itself cannot be used here: its
    np.min over two CUDA tensors (opt.py:463) does not run on a GPU."""
    import torch
    from optwboundeigenval_b200 import zoo
    zmodel, _ = zoo.build(kind)
    ref = _reference_objects(kind, zmodel, True)
    if ref is None:
        raise RuntimeError("no reference checkout on this machine")
    opt, model, loss = ref
    x, y = zoo.synthetic_batch(kind, batch)
    op = opt.HVPOperator(model, [x, y], loss, use_gpu=True)
    P = sum(p.numel() for p in model.parameters())
    v = torch.full((P,), P ** -0.5, dtype=torch.float64, device="cuda")
    t0 = 0.0
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            t0 = time.time()
        w = op.Hv(v, storedGrad=True)
        v = w / torch.norm(w)
    torch.cuda.synchronize()
    dt = time.time() - t0
    model.cpu()
    return steps / dt, dt


def reference_iter_body_rate(kind, batch, n_steps=2, k=20, use_gpu=False):
    """regularised steps/sec of the UNMODIFIED reference's iter() minibatch body (opt.py:608-699): comp_g with
    exactly k power iterations, vGHv (K = 0: penalty active), assembly, SGD step."""
    import torch
    from optwboundeigenval_b200 import zoo
    zmodel, _ = zoo.build(kind)
    ref = _reference_objects(kind, zmodel, use_gpu)
    if ref is None:
        return None
    opt, model, loss = ref
    xs, ys = zip(*[zoo.synthetic_batch(kind, batch, seed=zoo.SEED + j) for j in range(n_steps + 1)])
    x, y = torch.cat(xs), torch.cat(ys)
    cwd = os.getcwd()
    import tempfile
    tmp = tempfile.mkdtemp(prefix="b2s_ref_")
    os.makedirs(os.path.join(tmp, "logs"))
    os.chdir(tmp)
    stamps = []
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            o = opt.OptWBoundEignVal(model, loss, torch.optim.SGD(model.parameters(), lr=1e-4), mu=0.01, K=0,
                                     pow_iter_eps=0.0, max_pow_iter=k, use_gpu=use_gpu, ignore_bad_vals=False,
                                     header="bench_iter", batch_size=batch)
            real_g = o.comp_g

            def comp_g(data):                          # stamps the start of every minibatch body
                stamps.append(time.time())
                return real_g(data)
            o.comp_g = comp_g
            o.comp_f = lambda *a, **kw: (0.0, None)    # the epoch-end evaluation is not part of a step
            o.dataloader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=batch)
            o.iter()
    finally:
        os.chdir(cwd)
    dt = stamps[n_steps + 1] - stamps[1]               # bodies 1..n_steps; body 0 is the warm-up, the last stamp
    n = n_steps                                        # is the end-of-epoch comp_g(rdata) that follows them
    return {"value": n / dt, "unit": "regularized steps/s", "steps": n, "seconds": dt, "hvp_per_step": k,
            "cores": os.cpu_count() or 1, "kind": "reference",
            "sample": "%d minibatch bodies of the unmodified iter() after one warm-up body" % n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from optwboundeigenval_b200 import zoo
    kind, batch = args.config, args.batch or zoo.CONFIGS[args.config][3]
    val, dt, how = reference_comp_rho_rate(kind, batch, args.warmup, args.steps)
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(kind, batch, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": how,
                             "sample": "iterations %d..%d of comp_rho's loop on one %d-image minibatch (%s, nested "
                                       "torch.autograd on CPU, torch threads = %d)" % (
                                           args.warmup, args.warmup + args.steps - 1, batch,
                                           "unmodified opt.OptWBoundEignVal.comp_rho + opt.HVPOperator" if how == "reference"
                                           else "oracle/autograd_oracle.py", cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _config(kind, batch, world):
    names = {"cifar_densenet": "params/cifar10_DenseNet_mu0_01_K10.py: DenseNet3(40,12), 32x32 synthetic images",
             "usps": "params/usps_CNN_mu0_01_K0.py: USPS CNN, 16x16 synthetic digits",
             "forest": "params/forest_best.py: covertype MLP 54-20-20-20-7",
             "chest_vgg": "params/chestxray_mu0_001_K0_vgg.py: VGG16-bn chest model, 224x224 synthetic",
             "chest_densenet121": "params/chestxray_best_reg.py: DenseNet121 chest model, 224x224 synthetic"}
    return {"workload": names[kind], "batch_per_gpu": batch, "global_batch": batch * world,
            "unit_note": "one HVP = Hv of one %d-image minibatch (weak scaling: every GPU holds its own minibatch shard "
                         "of a global batch %d with synced BatchNorm; N GPUs complete N minibatch equivalents per step)"
                         % (batch, batch * world),
            "parallelism": "dp%d (minibatch sharded, synced BatchNorm sums, one all-reduce of the P-vector per HVP)" % world
            if world > 1 else "single GPU",
            "l2": "no flush: the HVP streams its value/tangent/adjoint caches, working set far above the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------------
def _golden_for_bench(kind, batch):
    """the reference's own outputs for the benched shape (tests/golden, generated by oracle/make_golden.py)"""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", kind + ".npz")
    if not os.path.exists(path):
        return None
    g = np.load(path, allow_pickle=False)
    if "x" not in g or g["x"].shape[0] != batch:
        return None
    return g


def per_config_table(args, skip):
    """device-resident HVP/s of the other BASELINE configs on this GPU (same kernels, same loop)"""
    import numpy as np
    import torch
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator, clear_plans
    rows = {}
    for kind in ("forest", "usps", "chest_densenet121", "chest_vgg"):
        if kind == skip:
            continue
        try:
            model, loss = zoo.build(kind)
            model.train()
            batch = zoo.CONFIGS[kind][3]
            x, y = zoo.synthetic_batch(kind, batch)
            op = B200HVPOperator(model, [x, y], loss)
            P = sum(p.numel() for p in model.parameters())
            v0 = torch.from_numpy(np.ones(P) / np.sqrt(P)).cuda()
            n = 200 if P < 100000 else 10
            op.power_iterate(v0, 0.0, 3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = op.power_iterate(v0, 0.0, n)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            rows[kind] = {"hvp_per_s": 1e3 / ms, "ms_per_step": ms, "batch": batch, "P": P, "steps": n,
                          "lambda": out.lam}
            del op, model
            clear_plans()
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001
            rows[kind] = {"error": repr(e)}
    return rows


def run_b200(args):
    # stdout carries exactly ONE JSON line (rank 0): everything else that writes to file descriptor 1 while the bench runs
    # (NCCL's "NCCL version ..." banner, library warnings) is sent to stderr, the line itself goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    from optwboundeigenval_b200 import _lib, zoo
    from optwboundeigenval_b200 import hvp_operator
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    kind = args.config
    batch = args.batch or zoo.CONFIGS[kind][3]
    warmup = max(args.warmup, 3)
    model, loss = zoo.build(kind)            # same seed on every rank: replicated parameters
    gold = _golden_for_bench(kind, batch)
    if gold is not None:                     # the reference's own weights for this shape: parity is checked below
        sd = model.state_dict()
        j, new = 0, {}
        for k, t in sd.items():
            n = t.numel()
            new[k] = torch.from_numpy(np.asarray(gold["state0"][j:j + n])).to(t.dtype).view(t.shape)
            j += n
        model.load_state_dict(new)
    model.train()
    x, y = zoo.synthetic_batch(kind, batch, seed=zoo.SEED + 1000 * rank)     # each rank its own shard
    op = B200HVPOperator(model, [x, y], loss)
    op.async_host_vectors = True             # pinned host vectors are copied without a host-side wait (see its docstring)
    P = sum(p.numel() for p in model.parameters())
    v0 = torch.from_numpy(np.ones(P) / np.sqrt(P)).cuda()
    hv_first = op.Hv(v0, storedGrad=True)    # base pass + first HVP: plan, workspaces, graph capture
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of the benched shape -------------------------------------------------------------------------
    parity = {}

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

    if world == 1 and gold is not None and np.array_equal(gold["x"], x.numpy()):
        tail = sum(q.numel() for q in list(model.parameters())[-2:])      # classifier weight + bias
        g_gpu = op.stored_grad.cpu().numpy()
        parity = {"against": "tests/golden/%s.npz (unmodified reference, CPU fp32, batch %d)" % (kind, batch),
                  "hv_rel_err": rel(hv_first.cpu().numpy(), gold["hv_v0"]),
                  "grad_rel_err": rel(g_gpu, gold["grad"]),
                  "grad_classifier_rel_err": rel(g_gpu[-tail:], gold["grad"][-tail:]),
                  "loss_rel_err": abs(float(op.loss_value) - float(gold["loss"])) / abs(float(gold["loss"])),
                  "tolerance": 1e-4,
                  "note": "whole-vector errors above the tolerance are ReLU decisions on pre-activations within fp32 rounding of "
                          "zero (two fp32 implementations with different summation orders decide them differently; one flipped "
                          "decision moves the adjoint of its layer by ~1/sqrt(#elements)): the loss and the classifier gradient, "
                          "which are continuous in those decisions, agree to fp32 rounding, and tests/test_gpu_parity.py::"
                          "test_grad_hv_vghv_match_reference_golden proves with the fp64 jet oracle conditioned on the GPU's "
                          "decisions that the GPU vectors are reproduced at rtol 1e-4 and that every differing decision is "
                          "fp32-ambiguous"}
    elif world > 1:
        # sharded HVP (synced BatchNorm sums + all-reduce) against the single-GPU HVP of the concatenated batch
        xs = [torch.empty_like(x).cuda() for _ in range(world)]
        ys = [torch.empty_like(y).cuda() for _ in range(world)]
        dist.all_gather(xs, x.cuda())
        dist.all_gather(ys, y.cuda())
        if rank == 0:
            # a plan of its own, outside the per-model cache and without a communicator (the cached sharded plan must stay
            # the one every rank uses); the BatchNorm running statistics its base pass updates are put back
            buffers = [t.clone() for t in model.buffers()]
            one = hvp_operator.SpectralPlan(model, loss, tuple(x.shape[1:]), batch * world, torch.device("cuda", local))
            g_one, _ = one.base_pass(hvp_operator.flat_parameters(model), torch.cat(xs), torch.cat(ys))
            hv_one = one.hv(v0)
            parity = {"against": "single-GPU HVP of the concatenated %d-image batch" % (batch * world),
                      "hv_rel_err": rel(hv_first.cpu().numpy(), hv_one.cpu().numpy()),
                      "grad_rel_err": rel(op.stored_grad.cpu().numpy(), g_one.cpu().numpy()),
                      "tolerance": 1e-4}
            del one
            with torch.no_grad():
                for t, saved in zip(model.buffers(), buffers):
                    t.copy_(saved)
        barrier()

    # ---- device-resident loop -------------------------------------------------------------------
    op.power_iterate(v0, 0.0, warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    out = op.power_iterate(v0, 0.0, args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert out.iters == args.steps - 1, "the timed loop must run exactly --steps iterations"
    if gold is not None and parity and "rho1_traj" in gold and world == 1:
        m = min(len(gold["rho1_traj"]), args.steps)
        traj = op.power_iterate(v0, 0.0, m, want_trajectory=True).trajectory
        parity["lambda_rel_err_first_iterations"] = float(np.max(np.abs(traj[:m, 1] - gold["rho1_traj"][:m, 1])
                                                                 / np.abs(gold["rho1_traj"][:m, 1])))
        parity["lambda_tolerance"] = 1e-3

    # ---- end to end through the reference-facing operator call, host vectors ----------------------
    v_host = torch.from_numpy(np.ones(P) / np.sqrt(P)).pin_memory()      # this step's input: pinned host memory
    r_host = torch.empty(P, dtype=torch.float64).pin_memory()            # this step's result, read back every step
    cur = torch.cuda.current_stream()
    vh, rh = v_host.numpy(), r_host.numpy()

    def e2e_step():
        r = op.Hv(v_host, storedGrad=True)          # 8P bytes host -> device inside the call (async, pinned)
        r_host.copy_(r, non_blocking=True)          # 8P bytes device -> host
        cur.synchronize()
        np.multiply(rh, 1.0 / np.sqrt(np.dot(rh, rh)), out=vh)          # next input depends on this output

    # The host-side normalisation between two calls is 176 k doubles: a threaded BLAS parks its workers during the
    # 2.3 ms the host waits for the device and pays their wake-up on every np.dot (measured: p95 of 20 ms against a
    # median of 40 us, the 212 .. 394 HVP/s spread of this figure between boxes), so the dot runs on the calling thread.
    try:
        from threadpoolctl import threadpool_limits
        blas_one = threadpool_limits(limits=1, user_api="blas")
    except Exception:                                   # threadpoolctl missing: keep numpy's default
        blas_one = None
    e2e_ms = []
    try:
        for _ in range(warmup):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            t_step = time.perf_counter()
            e2e_step()
            e2e_ms.append(1e3 * (time.perf_counter() - t_step))
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
    finally:
        if blas_one is not None:
            blas_one.restore_original_limits()

    # ---- second headline metric: regularised steps/sec ---------------------------------------------
    # one step = the minibatch body of iter() (opt.py:608-699): host batch -> device, new operator, base pass,
    # k = 20 power iterations (max_pow_iter = 20, pow_iter_eps = 0: pinned, as SURVEY 8d prescribes), penalty
    # gradient vGHv (K = 0 so the penalty is always active), fused step assembly + update.
    reg = None
    if not args.no_reg:
        try:
            from optwboundeigenval_b200.spectral import SpectralState
            st = SpectralState(model, loss, mu=0.01, K=0.0, pow_iter_eps=0.0, max_pow_iter=20, ignore_bad_vals=False)
            opt_sgd = torch.optim.SGD(model.parameters(), lr=1e-4)
            xh, yh = x.pin_memory(), y.pin_memory()
            n_reg = max(3, min(args.steps, 10))

            def reg_step():
                with contextlib.redirect_stdout(sys.stderr):          # comp_rho prints the reference's warnings
                    st.regularized_step([xh.to("cuda", non_blocking=True), yh.to("cuda", non_blocking=True)], opt_sgd)

            for _ in range(2):
                reg_step()
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(n_reg):
                reg_step()
            g1.record()
            barrier()
            ms_reg = g0.elapsed_time(g1) / n_reg
            tr = torch.tensor([ms_reg], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tr, op=dist.ReduceOp.MAX)
            ms_reg = float(tr[0])
            reg = {"value": world * 1e3 / ms_reg, "unit": "regularized steps/s", "ms_per_step": ms_reg,
                   "steps": n_reg, "hvp_per_step": 20, "penalty": "mu=0.01, K=0 (active every step): base pass + 20 HVPs + vGHv + "
                   "fused step assembly + SGD update; batch copied from pinned host memory every step",
                   "h2d_bytes_per_step": int(x.numel() * 4 + y.numel() * 8), "rho": float(st.rho)}
        except Exception as e:   # noqa: BLE001  (the headline line must survive a failure of this extra leg)
            reg = {"error": repr(e)}

    # ---- strong scaling: the config's batch as written (32 images) split over the N GPUs -----------------------------
    strong = None
    if world > 1 and batch % world == 0:
        per = batch // world
        xg, yg = zoo.synthetic_batch(kind, batch)                       # the same global batch on every rank
        op_s = B200HVPOperator(model, [xg[rank * per:(rank + 1) * per], yg[rank * per:(rank + 1) * per]], loss)
        op_s.Hv(v0, storedGrad=True)
        op_s.power_iterate(v0, 0.0, warmup)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        op_s.power_iterate(v0, 0.0, args.steps)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        strong = {"value": args.steps / (float(ts[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(ts[0]) / args.steps,
                  "global_batch": batch, "batch_per_gpu": per, "scaling": "strong",
                  "note": "the minibatch of the config as written split over the GPUs: %d images per GPU, the 78 synced "
                          "BatchNorm exchanges per HVP dominate (SURVEY 8e)" % per}
        op.Hv(v0, storedGrad=True)                                      # the weak-scaling shard is the cached base pass again

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel of the HVP pass (CUDA events around every launch, on the plan's stream) ----
    # conv_fwd and conv_dgrad are template instances of ONE kernel (conv_tma_kernel): they are one row here.
    peaks = _peaks()
    # every rank runs the profiled pass (it contains the pass's collectives); rank 0 reports
    prof = op.plan.profile(1, reps=3)
    line = None
    if rank == 0:
        fam = {}
        for r in prof:
            key = "conv_tma_kernel (fwd + dgrad instances)" if r["name"] in ("conv_fwd", "conv_dgrad") else r["name"]
            f = fam.setdefault(key, {"name": key, "ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            for k in ("ms", "flops", "bytes", "launches"):
                f[k] += r[k]
        top = max(fam.values(), key=lambda r: r["ms"])
        tot_ms = sum(r["ms"] for r in prof)
        is_gemm = top["name"].startswith("conv")
        try:
            tf32_peak, tf32_src = measure_tf32_peak(), "measured in this run: torch.matmul 8192^3 with allow_tf32, best of 10"
        except Exception as e:   # noqa: BLE001
            tf32_peak, tf32_src = peaks["bf16_tflops"] / 2.0, "fallback bf16/2 (%r)" % (e,)
        ai = top["flops"] / max(top["bytes"], 1.0)         # algorithmic FLOP per algorithmic byte
        balance = tf32_peak * 1e12 / (peaks["hbm_gbs"] * 1e9)
        gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        tfl = top["flops"] / (top["ms"] * 1e-3) / 1e12
        if is_gemm and ai >= balance:
            roof = {"bound": "tensor", "achieved": tfl, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": tfl / tf32_peak, "traffic": None}
        else:
            roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "traffic": None}
        ncu, ncu_path = _ncu_summary()
        if ncu and top["name"] in ncu:                # dram__bytes_read + dram__bytes_write per launch, ncu --set full capture
            roof["traffic"] = ncu[top["name"]].get("dram_bytes_per_launch")
            roof["traffic_source"] = ncu_path + ": " + ncu[top["name"]].get("capture", "")
        roof.update({"kernel": top["name"], "launches_per_step": top["launches"],
                     "ms_per_step_in_kernel": top["ms"], "share_of_step_kernel_time": top["ms"] / tot_ms,
                     "algorithmic_bytes_per_step": top["bytes"], "algorithmic_flops_per_step": top["flops"],
                     "arithmetic_intensity_flop_per_byte": ai, "machine_balance_flop_per_byte_tf32": balance,
                     "tensor_side": {"achieved_tflops": tfl, "peak_tf32_tflops": tf32_peak, "peak_source": tf32_src,
                                     "frac_of_tf32_peak": tfl / tf32_peak,
                                     "note": "3xTF32 emulation passes are not counted as algorithmic FLOPs"},
                     "peak_source": peaks["source"] + " (MEASURED_PEAKS.json burst copy bandwidth)",
                     "note": "thin layers (12..48 output channels): algorithmic intensity below the TF32 machine balance, "
                             "so the HBM roofline bounds the kernel; event time includes ~2 us of launch gap per launch"})
        cpu = cpu_reg = yard = None
        if world == 1 and not args.no_cpu:
            try:
                n_cpu = 12 if kind in ("cifar_densenet", "usps", "forest") else 2
                rate, dt, how = reference_comp_rho_rate(kind, batch, 2, n_cpu)
                cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": how,
                       "sample": "%d iterations of comp_rho's loop on the same %d-image minibatch in %.1f s after 2 warm-up "
                                 "iterations (%s)" % (n_cpu, batch, dt, "unmodified reference from baseline/_ref"
                                                      if how == "reference" else "oracle/autograd_oracle.py")}
            except Exception as e:   # noqa: BLE001
                cpu = {"error": repr(e)}
            if reg is not None and "error" not in reg and kind in ("cifar_densenet", "usps", "forest"):
                try:
                    cpu_reg = reference_iter_body_rate(kind, batch)
                except Exception as e:   # noqa: BLE001
                    cpu_reg = {"error": repr(e)}
            if not args.no_yardstick:
                # the reference's own autograd path on THIS GPU (its use_gpu=True branch: cuBLAS / cuDNN double backward,
                # torch defaults: cudnn.allow_tf32 = True, cuda.matmul.allow_tf32 = False) -- context, not the target
                try:
                    rate, dt = reference_gpu_hv_rate(kind, batch, 3, 10)
                    yard = {"value": rate, "unit": UNIT, "kind": "reference", "device": torch.cuda.get_device_name(),
                            "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
                            "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
                            "sample": "10 Hv(v, storedGrad=True) calls of the unmodified opt.HVPOperator(use_gpu=True) after 3 warm-up calls"}
                except Exception as e:   # noqa: BLE001
                    yard = {"error": repr(e)[:300]}
        vec = None
        if world == 1 and not args.no_vec:
            try:
                vec = vector_roofline(peaks)
            except Exception as e:   # noqa: BLE001
                vec = {"error": str(e)}
        table = None
        if world == 1 and not args.no_table:
            table = per_config_table(args, kind)
        if reg is not None and cpu_reg is not None:
            reg["cpu_reference"] = cpu_reg
        value = world * args.steps / (ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": _config(kind, batch, world),
                "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": 8 * P, "d2h_bytes_per_step": 8 * P,
                        "host_step_ms_median": float(np.median(e2e_ms)), "host_step_ms_max": float(np.max(e2e_ms)),
                        "vs_device_resident": ms / ms_e2e},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": roof,
                "roofline_vector_kernels": vec,
                "regularized_step": reg,
                "cpu_baseline": cpu,
                "gpu_autograd_yardstick": yard,
                "per_config": table,
                "parity": parity or None,
                "strong_scaling": strong,
                "kernel_profile_ms": {r["name"]: round(r["ms"], 4) for r in prof},
                "lambda_max": out.lam}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cifar_densenet")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-vec", action="store_true", help="skip the vector-kernel HBM roofline leg")
    ap.add_argument("--no-reg", action="store_true", help="skip the regularised-step leg")
    ap.add_argument("--no-yardstick", action="store_true", help="skip the reference-on-GPU (stock autograd) leg")
    ap.add_argument("--no-table", action="store_true", help="skip the per-config table")
    args = ap.parse_args()
    if args.steps < 1 or args.warmup < 0:
        raise SystemExit("bench.py: --steps must be >= 1 and --warmup >= 0")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
