"""Dev tool: executed thread-instructions per CUDA source line for one kernel of an .ncu-rep, by joining the report's
SASS page (per-instruction execution counts) with nvdisasm's line info for the same kernel in the built library.
usage: python tools/line_profile.py <report.ncu-rep> <strict|fast> <kernel-substring> <output pixels> [min instr/px]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, mode, pattern, px = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
thresh = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "openvr_fsr_b200/libovrfsr.so")], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f"kernels_{mode}.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(page.split("\n")))
kname = rows[0][1]
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 5]
ci, si = hdr.index("Thread Instructions Executed"), hdr.index("Source")
op = lambda s: re.sub(r"^@!?U?P\d+\s+", "", s.strip()).split()[0].split(".")[0]
want = [op(r[si]) for r in data]
best = None
starts = [i for i, l in enumerate(sass) if l.startswith(".text.") and pattern in l]
for st in starts:
    en = next(i for i in range(st + 1, len(sass)) if sass[i].lstrip().startswith(".section"))
    cur, ins = None, []
    for l in sass[st:en]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(2).strip(), cur))
    if len(ins) == len(want) and all(op(a[0]) == b for a, b in zip(ins, want)):
        best = (sass[st], ins)
        break
if not best:
    sys.exit(f"no function matching '{pattern}' has the report's instruction sequence ({len(want)} instr; kernel {kname}); rebuild differs from the profiled build")
print("kernel:", kname, "\nmatched:", best[0])
agg = collections.Counter()
for (s, loc), r in zip(best[1], data):
    agg[loc] += int(r[ci])
print("total thread-instr per px: %.1f" % (sum(agg.values()) / px))
src = {}
for loc, c in sorted(agg.items(), key=lambda kv: kv[0] or ("", 0)):
    if c / px < thresh or not loc:
        continue
    f, l = loc
    if f not in src:
        p = os.path.join(root, "openvr_fsr_b200/csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:110] if l - 1 < len(src[f]) else ""
    print(f"{f}:{l:4d} {c / px:7.1f}  {text}")
