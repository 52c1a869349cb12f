"""profiles/sass_summary.txt: SASS mnemonic counts per kernel of libd3feat_b200.so (`python tools/sass_summary.py`).
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy),
UTMALDG / UTMASTG = TMA tensor copies, SYNCS = mbarrier, HMMA = legacy mma.sync (B200_PROFILING.md)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "d3feat", "pytorch_b200", "libd3feat_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
pat = {"UTCHMMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA", "LDTM (tcgen05.ld)": r"\bLDTM", "STTM (tcgen05.st)": r"\bSTTM",
       "UTCBAR (tcgen05.commit)": r"\bUTCBAR", "UBLKCP (cp.async.bulk, TMA)": r"\bUBLKCP",
       "UTMALDG/UTMASTG (TMA tensor)": r"\bUTMA(LDG|STG)", "SYNCS (mbarrier)": r"\bSYNCS", "HMMA (mma.sync)": r"\bHMMA",
       "RED/ATOM": r"\b(RED|ATOMG|ATOM)\b", "LDG.E.128": r"LDG\.E\.128", "LDGSTS": r"\bLDGSTS"}
rows, tot = [], collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    c = {k: len(re.findall(p, f)) for k, p in pat.items()}
    for k, v in c.items():
        tot[k] += v
    if any(c[k] for k in list(pat)[:8]):
        rows.append((name, c))
out = ["# SASS evidence (round 2) -- `cuobjdump -sass d3feat/pytorch_b200/libd3feat_b200.so`, mnemonic counts per kernel", "",
       "Library built by `__graft_entry__.build()` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a -lineinfo -O3`).",
       "Mnemonics per B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit,",
       "UBLKCP = cp.async.bulk (TMA bulk copy), UTMALDG / UTMASTG = TMA tensor copies, SYNCS = mbarrier ops, HMMA = legacy mma.sync.",
       "", "## whole library", ""]
out += ["%-32s %6d" % (k, tot[k]) for k in pat]
out += ["", "## kernels that use the tensor cores / tensor memory / TMA / mbarriers", ""]
for name, c in sorted(rows, key=lambda r: r[0]):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    out.append(d[:150])
    out.append("    " + ", ".join("%s %d" % (k.split(" ")[0], v) for k, v in c.items() if v))
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[6:20]))
