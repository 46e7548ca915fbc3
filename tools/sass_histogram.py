"""Opcode histogram of every kernel in libnomad_b200.so (cuobjdump -sass): the evidence that the hot kernels are
tcgen05 / TMEM / TMA code (UTCHMMA / UTCQMMA, LDTM / STTM, UTMALDG / UTMASTG) and where mma.sync (HMMA) remains.

    python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nomad_b200/csrc/libnomad_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "IMMA", "DFMA", "FFMA",
       "HFMA2", "MUFU", "LDG", "STG", "LDS", "STS", "LDSM", "ATOM", "RED", "BAR", "SHFL", "LDGSTS"]
cur, hist, order = None, {}, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op, mods = m.group(1), m.group(2)
        hist[cur][op] += 1
        if op in ("UTCHMMA", "UTCQMMA") and ".2CTA" in mods:
            hist[cur][op + ".2CTA"] += 1
        if op == "HMMA":
            hist[cur]["HMMA" + mods.split(".F32")[0]] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"# {lib}: {len(order)} kernels")
for name, dn in zip(order, demangle):
    h = hist[name]
    tot.update(h)
    short = re.sub(r"\(.*", "", dn)[:110]
    keys = [k for k in h if any(k.startswith(p) for p in KEY)]
    body = " ".join(f"{k}={h[k]}" for k in sorted(keys, key=lambda k: (KEY.index(next(p for p in KEY if k.startswith(p))), k)))
    print(f"{short}\n    instr={sum(v for k, v in h.items() if '.' not in k)} {body}")
print("# library totals:", " ".join(f"{k}={tot[k]}" for k in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "DFMA")))
