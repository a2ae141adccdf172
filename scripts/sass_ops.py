"""Per-kernel SASS opcode histogram of the shipped library (evidence that the hot kernels are tcgen05 / TMEM / TMA code):
   python scripts/sass_ops.py > profiles/r02_sass_ops.txt        (runs anywhere cuobjdump is installed; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "raft_b200", "libraft_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDTM", "STTM", "UTCATOM", "SYNCS", "FFMA2", "FADD2", "FMUL2",
        "FFMA", "FMNMX3", "FMNMX", "HMMA", "LDS", "STS", "LDG", "STG", "ATOMG", "REDG", "RED", "BAR", "MUFU", "DFMA"]
fn, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        op = m.group(1)
        hist[fn][op] += 1
        hist[fn]["_total"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(so, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass; tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR,")
print("# tcgen05.ld -> LDTM, cp.async.bulk.tensor load/store -> UTMALDG/UTMASTG, cp.async.bulk -> UBLKCP, packed fp32 -> FFMA2/FADD2)")
tot = collections.Counter()
for (fn_, h), name in zip(hist.items(), demangle):
    short = re.sub(r"\(.*", "", name).replace("void b2d::", "b2d::")
    parts = [f"{k}={h[k]}" for k in KEYS if h[k]]
    print(f"{short:70s} instrs={h['_total']:6d}  " + " ".join(parts))
    tot.update(h)
print("TOTAL " + " ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
