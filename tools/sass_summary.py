"""Per-kernel SASS opcode summary of the shipped liblkm.so (cuobjdump -sass): the tensor-core / TMA / FP64-tensor opcodes that
prove which pipe a kernel uses.  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "linfa_b200", "liblkm.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "DMMA", "HMMA", "LDGSTS", "SYNCS", "FFMA2", "FFMA", "DFMA", "ATOM", "RED"]
fn, counts, order = None, {}, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        order.append(fn)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[fn][k] += 1
                break
demangle = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
print(f"# {os.path.basename(lib)}: SASS opcode counts per kernel (sm_100a).  UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load,")
print("# UBLKCP = bulk copy, UTCBAR = tcgen05.commit, DMMA = FP64 tensor-core mma, LDGSTS = cp.async, SYNCS = mbarrier, FFMA2 = packed FP32 FMA")
print(f"{'kernel':84s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
for f, name in zip(order, demangle):
    c = counts[f]
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("lkm::", "")
    print(f"{short[:84]:84s} {c['total']:6d} " + " ".join(f"{c[k]:7d}" if c[k] else f"{'.':>7s}" for k in KEYS))
