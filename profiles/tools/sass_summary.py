#!/usr/bin/env python
"""Instruction mix of the hot kernels from `cuobjdump -sass` of the built library: counts of the mnemonics DESIGN.md names
(packed-byte SAD, 16-bit-lane min/max/add, dot product, warp reductions, TMA, mbarrier) and a short excerpt around the first use
of each. usage: sass_summary.py libfh264_b200.so out.md"""
import re, subprocess, sys
from collections import Counter

lib, out_md = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line.strip())
KERNELS = ["k_stage3", "k_stage2", "k_spec", "k_skipspec", "k_phase_b_warp", "k_phase_bPK", "k_phase_c", "k_interp", "k_features", "k_tile_index", "k_scene_sad", "k_intra", "k_cavlc_code"]
WATCH = ["VABSDIFF4", "VIADDMNMX", "VIMNMX", "VIADD", "IDP", "REDUX", "UTMALDG", "SYNCS", "UBLKCP", "SHFL", "VOTE", "POPC", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED.", "BAR", "NANOSLEEP"]
def mnem(l):
    m = re.search(r"\*/\s+(@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", l)
    return m.group(2) if m else ""
with open(out_md, "w") as f:
    f.write("# SASS instruction mix of the built library (cuobjdump -sass, sm_100a)\n\n")
    f.write("Counts are STATIC instructions in the kernel's code (not executed counts). Columns: mnemonic prefix -> occurrences.\n\n")
    for k in KERNELS:
        names = [n for n in funcs if k in n]
        if not names:
            continue
        body = funcs[names[0]]
        ms = [mnem(l) for l in body]
        c = Counter()
        for m_ in ms:
            for w in WATCH:
                if m_.startswith(w):
                    c[w] += 1
        full = Counter(m_ for m_ in ms if any(m_.startswith(w) for w in ("VABSDIFF4", "VIADDMNMX", "VIMNMX", "IDP", "REDUX", "UTMALDG", "SYNCS", "VIADD")))
        f.write("## %s  (%d instructions)\n\n" % (names[0], len(body)))
        f.write(", ".join("%s %d" % (w, c[w]) for w in WATCH if c[w]) + "\n\n")
        f.write("exact forms: " + ", ".join("`%s` x%d" % (m_, n) for m_, n in full.most_common(14)) + "\n\n")
        shown = set()
        for w in ("UTMALDG", "VABSDIFF4", "VIADDMNMX", "IDP", "REDUX"):
            for i, m_ in enumerate(ms):
                if m_.startswith(w) and w not in shown:
                    shown.add(w)
                    f.write("first `%s`:\n```\n%s\n```\n" % (w, "\n".join(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l)[:150] for l in body[max(0, i - 3):i + 4])))
                    break
        f.write("\n")
print("wrote", out_md)
