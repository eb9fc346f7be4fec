"""Dev tool: executed thread-instructions per CUDA source line straight from an .ncu-rep captured with
--import-source on (no rebuild needed: uses the report's own source correlation).
usage: python tools/line_profile2.py <report.ncu-rep> <output pixels> [min instr/px] [launch-id]"""
import csv, io, subprocess, sys
rep, px = sys.argv[1], float(sys.argv[2])
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, tot, lines, kern = None, None, 0, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if kern and r[1] != kern:
            break
        kern = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ti = hdr.index("Thread Instructions Executed"); continue
    if hdr and r[0].isdigit():
        try:
            c = int(r[ti])
        except ValueError:
            continue
        tot += c
        lines.append((fname, int(r[0]), c, r[1].strip()[:120]))
print("kernel:", kern)
print("total thread-instr per px: %.1f" % (tot / px))
for f, l, c, s in lines:
    if c / px >= thresh:
        print(f"{f}:{l:4d} {c / px:7.1f}  {s}")
