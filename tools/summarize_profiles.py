"""Turns the ncu captures of tools/capture_profiles.sh (gpurun_out/<round>_*.ncu-rep, *_launches_bench.csv) into the
small tracked summaries under profiles/: per-kernel key metrics (CSV), the launch-list aggregate of one power-iteration
step (share of step per kernel) and profiles/<round>_ncu_top_kernel.json, which bench.py reads for roofline.traffic.
Runs in the build container (ncu -i needs no GPU)."""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
SRC = os.environ.get("B2S_PROFILE_SRC", os.path.join(ROOT, "gpurun_out"))
DST = os.environ.get("B2S_PROFILE_DST", os.path.join(ROOT, "profiles"))
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarize_rep(name):
    rep = os.path.join(SRC, "%s_%s.ncu-rep" % (R, name))
    if not os.path.exists(rep):
        return []
    hdr, units, rows = raw_page(rep)
    idx = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    out = []
    with open(os.path.join(DST, "%s_ncu_%s.csv" % (R, name)), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([k + (" [%s]" % units[i] if units[i] else "") for k, i in idx])
        for r in rows:
            w.writerow([r[i] for _, i in idx])
            out.append({k: (r[i], units[i]) for k, i in idx})
    return out


def launch_list():
    path = os.path.join(SRC, "%s_launches_bench.csv" % R)
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui, si = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Stream")
    rows = []
    for r in rd:
        full = re.sub(r"^void ", "", r[ki]).replace("b2s::", "")
        v = float(r[vi].replace(",", ""))
        rows.append((re.sub(r"[<(].*", "", full), v / 1000 if r[ui] == "ns" else v, r[si]))
    upd = [i for i, r in enumerate(rows) if r[0] == "pi_update_kernel"]
    if len(upd) < 4:
        return
    a, b = upd[-3], upd[-2]                       # one power-iteration step of the timed loop
    seg = rows[a + 1:b + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    streams = collections.defaultdict(float)
    for n, v, s in seg:
        agg[n][0] += 1
        agg[n][1] += v
        streams[s] += v
    tot = sum(v for _, v in agg.values())
    with open(os.path.join(DST, "%s_launches_one_step.csv" % R), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "launches_per_step", "device_us_per_step (ncu, cold cache, serialised)", "share_of_step"])
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([n, c, "%.1f" % v, "%.4f" % (v / tot)])
        w.writerow(["TOTAL", len(seg), "%.1f" % tot, "1.0"])
        for s, v in streams.items():
            w.writerow(["stream %s" % s, "", "%.1f" % v, "%.4f" % (v / tot)])
    # the full list is ~1 MB: keep the first base pass + three steps
    with open(os.path.join(DST, "%s_launches_bench_head.csv" % R), "w") as fh:
        fh.write("".join(lines[: upd[min(3, len(upd) - 1)] + 3]))
    return agg, tot


def main():
    os.makedirs(DST, exist_ok=True)
    agg = launch_list()
    conv = summarize_rep("conv_tma")
    wg = summarize_rep("conv_wgrad")
    summarize_rep("bn")
    vec = summarize_rep("vec")
    vgg = summarize_rep("vgg_conv_tc")
    vggw = summarize_rep("vgg_wgrad")
    summarize_rep("vgg_bn")
    def per_launch(rows):
        per = []
        for r in rows:
            rd = to_bytes(*r["dram__bytes_read.sum"])
            wr = to_bytes(*r["dram__bytes_write.sum"])
            per.append({"kernel": r["Kernel Name"][0][:80], "grid": r["Grid Size"][0], "us": float(r["gpu__time_duration.sum"][0]),
                        "dram_bytes": rd + wr,
                        "tensor_pipe_active_pct": float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]),
                        "issue_active_pct": float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"][0])})
        return per

    js = {}
    for key, rows, what in (("conv_tma_kernel (fwd + dgrad instances)", conv, "conv_tma"), ("conv_wgrad", wg, "conv_wgrad"),
                            ("chest_vgg conv_tc_kernel", vgg, "vgg_conv_tc"), ("chest_vgg conv_wgrad", vggw, "vgg_wgrad")):
        if rows:
            per = per_launch(rows)
            js[key] = {"capture": "%s_%s.ncu-rep: ncu --set full --clock-control none, %d consecutive launches of the Hv pass "
                                  "(%s, eager launches)" % (R, what, len(per), "VGG16-bn, batch 4" if what.startswith("vgg") else "DenseNet3, batch 32"),
                       "dram_bytes_per_launch": sum(p["dram_bytes"] for p in per) / len(per), "launches": per}
    if js:
        with open(os.path.join(DST, "%s_ncu_top_kernel.json" % R), "w") as fh:
            json.dump(js, fh, indent=1)
    if vec:
        for r in vec:
            print(r["Kernel Name"][0][:40], r["gpu__time_duration.sum"], r["dram__bytes_read.sum"], r["dram__bytes_write.sum"],
                  r["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"])
    for f in ("%s_bench_plain.json" % R,):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
    print(sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
