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
        BatchNorm statistics, one NCCL all-reduce of the P-vector per HVP); value counts
        32-image-minibatch HVP equivalents: N * K / time.
e2e     the same HVP through the reference-facing call B200HVPOperator.Hv(vec) with a HOST vector:
        per step 8P bytes host->device (pinned) and 8P bytes device->host are inside the timed region.
roofline / cpu_baseline: see DESIGN.md "Measurement".
--impl reference times the CPU restatement of the reference's algorithm (oracle/, nested torch.autograd
on all host cores) on the same config; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


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
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)"""
    path = os.path.join(ROOT, "profiles", "r1_ncu_top_kernel.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return None


def vector_roofline(peaks):
    """HBM roofline of the eigen-iteration's two fused vector kernels at P = 2^26 (DenseNet3's P = 176 122
    sits in L2, so the >= 90 % claim is measured at a DRAM-resident size; 52 algorithmic bytes per element)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_vec
    r = bench_vec.measure(1 << 26, iters=30, warmup=5)
    return {"kernel": "pi_dot_kernel + pi_update_kernel", "P": r["n"], "bound": "hbm", "achieved": r["GBps"],
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": r["GBps"] / peaks["hbm_gbs"],
            "ms_per_iteration": r["ms_per_iteration"], "algorithmic_bytes_per_iteration": r["bytes_per_iteration"]}


def cpu_hvp_rate(kind, batch, seconds_budget=20.0, min_calls=3):
    """HVPs/sec of the oracle port (the reference's nested-autograd algorithm) on all host cores."""
    import torch
    from optwboundeigenval_b200 import zoo
    from oracle import autograd_oracle as ao
    torch.set_num_threads(os.cpu_count() or 1)
    model, loss = zoo.build(kind)
    model.train()
    x, y = zoo.synthetic_batch(kind, batch)
    op = ao.AutogradSpectralOperator(model, [x, y], loss)
    P = sum(p.numel() for p in model.parameters())
    v = ao.start_vector(P)
    op.hv(v)                      # builds the graph (prepare_grad) + first double backward: warm-up
    op.hv(v)
    t0 = time.time()
    n = 0
    while n < min_calls or (time.time() - t0 < seconds_budget and n < 200):
        op.hv(v)
        n += 1
    dt = time.time() - t0
    return n / dt, n, dt


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from optwboundeigenval_b200 import zoo
    from oracle import autograd_oracle as ao
    kind, batch = args.config, args.batch or zoo.CONFIGS[args.config][3]
    torch.set_num_threads(os.cpu_count() or 1)
    model, loss = zoo.build(kind)
    model.train()
    x, y = zoo.synthetic_batch(kind, batch)
    op = ao.AutogradSpectralOperator(model, [x, y], loss)
    P = sum(p.numel() for p in model.parameters())
    v = ao.start_vector(P)
    for _ in range(max(args.warmup, 1)):
        op.hv(v)
    t0 = time.time()
    for _ in range(args.steps):
        w = op.hv(v)
        lam = float(torch.dot(w, v))
        v = w / torch.norm(w) if lam >= 0 else -w / torch.norm(w)
    dt = time.time() - t0
    val = args.steps / dt
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "HVPs/sec (power-iter lambda_max)", "value": val, "unit": "HVP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(kind, batch, 1),
            "cpu_baseline": {"value": val, "unit": "HVP/s", "cores": cores, "kind": "port",
                             "sample": "%d power-iteration HVPs of one %d-image minibatch (oracle/autograd_oracle.py, "
                                       "nested torch.autograd on CPU, torch threads = %d)" % (args.steps, batch, cores)},
            "e2e": {"value": val, "unit": "HVP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _config(kind, batch, world):
    names = {"cifar_densenet": "params/cifar10_DenseNet_mu0_01_K10.py: DenseNet3(40,12), 32x32 synthetic images",
             "usps": "params/usps_CNN_mu0_01_K0.py: USPS CNN, 16x16 synthetic digits",
             "forest": "params/forest_best.py: covertype MLP 54-20-20-20-7",
             "chest_vgg": "params/chestxray_mu0_001_K0_vgg.py: VGG16-bn chest model, 224x224 synthetic",
             "chest_densenet121": "params/chestxray_best_reg.py: DenseNet121 chest model, 224x224 synthetic"}
    return {"workload": names[kind], "batch_per_gpu": batch, "global_batch": batch * world,
            "parallelism": "dp%d (minibatch sharded, synced BatchNorm sums, one NCCL all-reduce of the P-vector per HVP)" % world
            if world > 1 else "single GPU",
            "l2": "no flush: the HVP streams its value/tangent/adjoint caches, working set far above the 126 MB L2"}


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
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    kind = args.config
    batch = args.batch or zoo.CONFIGS[kind][3]
    model, loss = zoo.build(kind)            # same seed on every rank: replicated parameters
    model.train()
    x, y = zoo.synthetic_batch(kind, batch, seed=zoo.SEED + 1000 * rank)     # each rank its own shard
    op = B200HVPOperator(model, [x, y], loss)
    P = sum(p.numel() for p in model.parameters())
    v0 = torch.from_numpy(np.ones(P) / np.sqrt(P)).cuda()
    op.Hv(v0, storedGrad=True)               # base pass + first HVP: plan, workspaces, graph capture
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop -------------------------------------------------------------------
    op.power_iterate(v0, 0.0, max(args.warmup, 3))
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

    # ---- end to end through the reference-facing operator call, host vectors ----------------------
    v_host = torch.from_numpy(np.ones(P) / np.sqrt(P)).pin_memory()      # this step's input: pinned host memory
    r_host = torch.empty(P, dtype=torch.float64).pin_memory()            # this step's result, read back every step
    cur = torch.cuda.current_stream()

    def e2e_step():
        r = op.Hv(v_host, storedGrad=True)          # 8P bytes host -> device inside the call
        r_host.copy_(r, non_blocking=True)          # 8P bytes device -> host
        cur.synchronize()
        a = r_host.numpy()
        np.divide(a, np.linalg.norm(a), out=v_host.numpy())             # next input depends on this output

    for _ in range(3):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- second headline metric: regularised steps/sec ---------------------------------------------
    # one step = the minibatch body of iter() (opt.py:608-699): host batch -> device, new operator, base pass,
    # k = 20 power iterations (max_pow_iter = 20, pow_iter_eps = 0: pinned, as SURVEY 8d prescribes), penalty
    # gradient vGHv (K = 0 so the penalty is always active), fused step assembly, SGD update.
    reg = None
    if not args.no_reg:
        try:
            import contextlib
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
            reg = {"value": world * 1e3 / ms_reg, "unit": "regularized steps/s (32-image minibatches)", "ms_per_step": ms_reg,
                   "steps": n_reg, "hvp_per_step": 20, "penalty": "mu=0.01, K=0 (active every step): base pass + 20 HVPs + vGHv + "
                   "fused step assembly + SGD update; batch copied from pinned host memory every step",
                   "h2d_bytes_per_step": int(x.numel() * 4 + y.numel() * 8), "rho": float(st.rho)}
        except Exception as e:   # noqa: BLE001  (the headline line must survive a failure of this extra leg)
            reg = {"error": repr(e)}

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel of the HVP pass (CUDA events around every launch, on the plan's stream) ----
    # conv_fwd and conv_dgrad are template instances of ONE kernel (conv_tma_kernel): they are one row here.
    peaks = _peaks()
    # every rank runs the profiled pass (it contains the pass's NCCL all-reduces); rank 0 reports
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
        tf32_peak = peaks["bf16_tflops"] / 2.0            # dense TF32 = half the dense bf16 rate on B200 (1.1 vs 2.25 PFLOP/s)
        ai = top["flops"] / max(top["bytes"], 1.0)         # algorithmic FLOP per algorithmic byte
        balance = tf32_peak * 1e12 / (peaks["hbm_gbs"] * 1e9)
        gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        tfl = top["flops"] / (top["ms"] * 1e-3) / 1e12
        if is_gemm and ai >= balance:
            roof = {"bound": "tensor", "achieved": tfl, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": tfl / peaks["bf16_tflops"], "traffic": None}
        else:
            roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "traffic": None}
        ncu = _ncu_summary()
        if ncu and top["name"] in ncu:                # dram__bytes_read + dram__bytes_write per launch, ncu --set full capture
            roof["traffic"] = ncu[top["name"]].get("dram_bytes_per_launch")
            roof["traffic_source"] = "profiles/r1_ncu_top_kernel.json: " + ncu[top["name"]].get("capture", "")
        roof.update({"kernel": top["name"], "launches_per_step": top["launches"],
                     "ms_per_step_in_kernel": top["ms"], "share_of_step_kernel_time": top["ms"] / tot_ms,
                     "algorithmic_bytes_per_step": top["bytes"], "algorithmic_flops_per_step": top["flops"],
                     "arithmetic_intensity_flop_per_byte": ai, "machine_balance_flop_per_byte_tf32": balance,
                     "tensor_side": {"achieved_tflops": tfl, "peak_bf16_tflops": peaks["bf16_tflops"],
                                     "frac_of_bf16_peak": tfl / peaks["bf16_tflops"],
                                     "note": "3xTF32 emulation passes are not counted as algorithmic FLOPs"},
                     "peak_source": peaks["source"] + " (MEASURED_PEAKS.json burst copy bandwidth / bf16 matmul)",
                     "note": "thin layers (12..48 output channels): algorithmic intensity below the TF32 machine balance, "
                             "so the HBM roofline bounds the kernel; event time includes ~2 us of launch gap per launch"})
        cpu = None
        if world == 1 and not args.no_cpu:
            rate, n, dt = cpu_hvp_rate(kind, batch)
            cpu = {"value": rate, "unit": "HVP/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": "%d Hv calls of the same %d-image minibatch in %.1f s (oracle/autograd_oracle.py)" % (n, batch, dt)}
        vec = None
        if world == 1 and not args.no_vec:
            try:
                vec = vector_roofline(peaks)
            except Exception as e:   # noqa: BLE001
                vec = {"error": str(e)}
        value = world * args.steps / (ms * 1e-3)
        line = {"metric": "HVPs/sec (power-iter lambda_max)", "value": value,
                "unit": "HVP/s (32-image minibatch equivalents)" if kind == "cifar_densenet" else "HVP/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": _config(kind, batch, world),
                "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "HVP/s",
                        "h2d_bytes_per_step": 8 * P, "d2h_bytes_per_step": 8 * P},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": roof,
                "roofline_vector_kernels": vec,
                "regularized_step": reg,
                "cpu_baseline": cpu,
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cifar_densenet")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-vec", action="store_true", help="skip the vector-kernel HBM roofline leg")
    ap.add_argument("--no-reg", action="store_true", help="skip the regularised-step leg")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 12)
        args.warmup = min(args.warmup, 2)
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
