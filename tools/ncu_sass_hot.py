"""Summarise `ncu --page source --csv --print-source sass`: instructions executed and stall samples per
code region (regions split at a given list of SASS opcodes / addresses), plus the hottest instructions."""
import csv, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; col = {n: i for i, n in enumerate(h)}
data = []
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[0] == "Address":
        break
    data.append(r)
tot_inst = sum(float(r[col["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(float(r[col["# Samples"]] or 0) for r in data)
print("total warp-instr", tot_inst, "samples", tot_samp)
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
# running windows of 64 instructions
W = int(sys.argv[2]) if len(sys.argv) > 2 else 48
print(f"{'idx':>5s} {'addr':>8s} {'inst%':>6s} {'samp%':>6s}  top stalls   first-op")
for s in range(0, len(data), W):
    blk = data[s:s + W]
    inst = sum(float(r[col["Instructions Executed"]] or 0) for r in blk)
    samp = sum(float(r[col["# Samples"]] or 0) for r in blk)
    if inst / tot_inst < 0.003 and samp / tot_samp < 0.003:
        continue
    st = {n: sum(float(r[col[n]] or 0) for r in blk) for n in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    ops = " ".join(r[col["Source"]].split()[0] for r in blk[:6])
    print(f"{s:5d} {blk[0][col['Address']][-5:]:>8s} {100*inst/tot_inst:6.2f} {100*samp/tot_samp:6.2f}  " +
          " ".join(f"{k[6:]}={int(v)}" for k, v in top) + "   " + ops)
print("hottest instructions by samples:")
for r in sorted(data, key=lambda r: -float(r[col["# Samples"]] or 0))[:25]:
    st = {n: float(r[col[n]] or 0) for n in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {r[col['Address']][-5:]} samp={r[col['# Samples']]:>5s} inst={r[col['Instructions Executed']]:>8s} {r[col['Source']][:70]:70s} " +
          " ".join(f"{k[6:]}={int(v)}" for k, v in top))
