"""Dev tool: static SASS structure of one kernel -- every backward branch (loop) with its body size and opcode mix.
usage: python tools/sass_loops.py <object-or-so> <mangled-kernel-substring> [min body size]"""
import re, subprocess, sys
from collections import Counter
obj, pat = sys.argv[1], sys.argv[2]
minsz = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for l in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(f"{name}: {len(ins)} instructions")
    addr2i = {a: i for i, (a, _) in enumerate(ins)}
    op = lambda s: re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
    for i, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", s)
        if m and int(m.group(1), 16) in addr2i and int(m.group(1), 16) <= a:
            j = addr2i[int(m.group(1), 16)]
            if i - j + 1 >= minsz:
                c = Counter(op(x) for _, x in ins[j:i + 1])
                print(f"  loop 0x{ins[j][0]:x}..0x{a:x}: {i - j + 1} instr: " + " ".join(f"{k}:{v}" for k, v in c.most_common(30)))
    c = Counter(op(x) for _, x in ins)
    print("  whole: " + " ".join(f"{k}:{v}" for k, v in c.most_common(40)))
