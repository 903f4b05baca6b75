"""Instruction-count summary of the in-tree library's SASS (cuobjdump -sass): per kernel, how many tcgen05 / TMA / TMEM
instructions it holds -- the evidence that the kernels are Blackwell-native (B200_PROFILING.md: UTCHMMA = tcgen05.mma,
UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit -> mbarrier).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "video-captioning-transformer_b200", "libvct_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "MUFU",
         "LDGSTS", "UBLKCP", "ELECT", "ACQBULK", "CCTL"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["/usr/local/cuda/bin/cu++filt", s], capture_output=True, text=True).stdout.strip()
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][w + ("." + op.split(".")[1] if w == "UTMALDG" and "." in op else "")] += 1
tot = collections.Counter()
print(f"# {os.path.basename(so)}: {len(per)} kernels, sm_100a SASS instruction counts (cuobjdump -sass)")
for fn, c in per.items():
    name = demangle(fn)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|vct::|void ", "", name)
    name = re.sub(r">\(.*", ">", name) if "<" in name else re.sub(r"\(.*", "", name)
    name = name.replace("(bool)", "").replace("(int)", "")
    keys = [k for k in c if k != "_total" and c[k]]
    if not any(k.startswith(("UTC", "UTMA", "LDTM", "STTM")) for k in keys):
        continue
    print(f"{name[:78]:78s} total {c['_total']:6d}  " + "  ".join(f"{k} {c[k]}" for k in sorted(keys) if k not in ("FFMA", "MUFU", "CCTL")))
    tot.update(c)
print("\n# sum over the tensor-core / TMA kernels above")
print("  ".join(f"{k} {v}" for k, v in sorted(tot.items()) if k != "_total"))
