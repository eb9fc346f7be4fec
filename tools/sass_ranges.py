"""Dev tool: executed thread-instructions per output pixel of an .ncu-rep, binned by SASS address ranges
(offsets from the first instruction).  usage: python tools/sass_ranges.py <rep> <px> <hex offset> [<hex offset> ...]"""
import csv, io, subprocess, sys
rep, px = sys.argv[1], float(sys.argv[2])
cuts = sorted(int(x, 16) for x in sys.argv[3:])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
iA, iS, iE = hdr.index('Address'), hdr.index('Source'), hdr.index('Thread Instructions Executed')
data = [(int(r[iA], 16), r[iS].strip(), int(r[iE])) for r in rows[2:] if len(r) >= len(hdr) and r[iA].startswith('0x')]
base = data[0][0]
bins = {}
for a, s, e in data:
    off = a - base
    k = max([c for c in cuts if c <= off], default=0)
    bins[k] = bins.get(k, 0) + e
for k in sorted(bins):
    print(f"from 0x{k:x}: {bins[k] / px:8.1f} thread-instr/px")
print(f"total {sum(bins.values()) / px:.1f}")
