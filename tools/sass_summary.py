"""Tensor-core / TMA / TMEM mnemonics per contraction-kernel instance of the built library (cuobjdump -sass; runs in the
build container, no GPU needed).  Writes profiles/<round>_sass_contraction_kernels.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
LIB = os.path.join(ROOT, "optwboundeigenval_b200", "libb200spectral.so")
KERNELS = ("conv_tma_kernel", "conv_tc_kernel", "conv_tc_wgrad_kernel", "conv_wgrad_tma_kernel")
COUNT = ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDG", "LDS", "STS", "FFMA", "RED", "ATOMG")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    names = dict(zip(re.findall(r"Function : (\S+)", sass), demangle))
    out = ["cuobjdump -sass optwboundeigenval_b200/libb200spectral.so (sm_100a), tensor-core / TMA / TMEM mnemonics per kernel instance",
           "UTCHMMA = tcgen05.mma (kind::tf32), UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk, LDTM/STTM = tcgen05.ld/st (TMEM), "
           "UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, RED/ATOMG = global atomics", ""]
    for block in sass.split("Function : ")[1:]:
        mangled = block.split("\n", 1)[0].strip()
        name = names.get(mangled, mangled)
        if not any(k + "<" in name for k in KERNELS):
            continue
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", block, flags=re.M)
        cnt = collections.Counter()
        variants = set()
        for op in ops:
            base = op.split(".")[0]
            if base in COUNT:
                cnt[base] += 1
                if base in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM"):
                    variants.add(op)
        out.append(re.sub(r"\(.*", "", name))
        out.append("   " + "  ".join("%s=%d" % (k, cnt[k]) for k in COUNT if cnt[k]))
        out.append("   variants: " + ", ".join(sorted(variants)))
    path = os.path.join(ROOT, "profiles", "%s_sass_contraction_kernels.txt" % R)
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print(path, len(out))


if __name__ == "__main__":
    main()
