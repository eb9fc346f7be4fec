"""Static SASS evidence for profiles/: per kernel of the built library, the instruction counts that show what the code is
made of -- TMA (UTMALDG / UTMASTG / UBLKCP), mbarrier (SYNCS), packed FP32x2 (FFMA2 / FMUL2 / FADD2), store widths
(STG.E / .64 / .128), shared-memory access widths, MUFU, barriers, branches -- plus registers from the resource usage.
usage: python tools/sass_summary.py [libovrfsr.so] > profiles/r2_sass_summary.txt"""
import re, subprocess, sys
from collections import Counter, OrderedDict
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "openvr_fsr_b200" / "libovrfsr.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for l in res.split("\n"):
    m = re.match(r"\s*Function (\S+):", l)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", l)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
COLS = OrderedDict([
    ("UTMALDG", r"^UTMALDG"), ("UTMASTG", r"^UTMASTG"), ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("UGETNEXTWORKID", r"^UGETNEXTWORKID"),
    ("FFMA2", r"^FFMA2"), ("FMUL2", r"^FMUL2"), ("FADD2", r"^FADD2"), ("FFMA", r"^FFMA\b"), ("FMUL", r"^FMUL\b"), ("FADD", r"^FADD\b"),
    ("STG.32", r"^STG\.E\s"), ("STG.64", r"^STG\.E\.64"), ("STG.128", r"^STG\.E\.128"),
    ("LDG.128", r"^LDG\.E\.128"), ("LDS.32", r"^LDS\s"), ("LDS.64", r"^LDS\.64"), ("LDS.128", r"^LDS\.128"), ("STS.128", r"^STS\.128"),
    ("MUFU", r"^MUFU"), ("FCHK", r"^FCHK"), ("BAR", r"^BAR"), ("BRA", r"^BRA"), ("BSSY", r"^BSSY"),
])
blocks = re.split(r"\n\s*Function : ", sass)
rows = []
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    ins = []
    for l in b.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            ins.append(re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()))
    c = Counter()
    for i in ins:
        for k, pat in COLS.items():
            if re.search(pat, i):
                c[k] += 1
    rows.append((name, len(ins), c))
print(f"# static SASS instruction counts per kernel of {Path(lib).name} (cuobjdump -sass; sm_100a only)")
print("# TMA loads = UTMALDG, TMA stores = UTMASTG, mbarrier = SYNCS, cluster launch control = UGETNEXTWORKID, packed FP32x2 = FFMA2 / FMUL2 / FADD2")
tot = Counter()
for name, n, c in sorted(rows, key=lambda r: r[0]):
    d = demangle(name)
    d = (d[: d.rfind(">(") + 1] if ">(" in d else d.split("(")[0]).replace("ovrfsr::", "").replace("void ", "")
    r = regs.get(name, (0, 0))
    print(f"{d}\n    instr {n:5d}  regs {r[0]:3d}  static smem {r[1]:6d} B   " + "  ".join(f"{k}:{v}" for k, v in c.items() if v))
    tot.update(c)
print("\nTOTAL over all kernels: " + "  ".join(f"{k}:{tot[k]}" for k in COLS if tot[k]))
