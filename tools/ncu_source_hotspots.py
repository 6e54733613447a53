"""Per-SASS-instruction view of one kernel of an .ncu-rep (ncu --set full --import-source on): instruction counts, sampled
stall reasons, and a roll-up by opcode class.  Usage: python tools/ncu_source_hotspots.py REP KERNEL_REGEX [top_n]"""
import csv, io, re, subprocess, sys
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], text=True)
blocks = out.split('"Kernel Name"')
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = tot_samples = 0
per_op, per_op_samples, stall_tot = Counter(), Counter(), Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].strip()
    op = sass.split()[0] if sass else "?"
    if op.startswith("@"):
        op = sass.split()[1]
    op = op.rstrip(";")
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    tot_inst += n; tot_samples += s
    key = re.split(r"[.\s]", op)[0]
    per_op[key] += n; per_op_samples[key] += s
    for c in stall_cols:
        stall_tot[c] += int(r[ix[c]] or 0)
    lines.append((s, n, sass, {c: int(r[ix[c]] or 0) for c in stall_cols}))
print(f"kernel {rows[0][1][:80]}: {tot_inst/1e6:.1f} M warp instructions, {tot_samples} samples")
print("stall samples:", ", ".join(f"{k[6:]} {v} ({100*v/max(1,sum(stall_tot.values())):.0f}%)" for k, v in stall_tot.most_common(8)))
print("by opcode (instructions M / share / samples share):")
for k, v in per_op.most_common(25):
    print(f"  {k:10s} {v/1e6:9.1f}  {100*v/tot_inst:5.1f}%   samples {100*per_op_samples[k]/max(1,tot_samples):5.1f}%")
print(f"top {topn} instructions by samples:")
for s, n, sass, st in sorted(lines, key=lambda t: -t[0])[:topn]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {s:6d} smp {n/1e6:8.2f} M  {sass[:70]:70s} {top[0][0][6:]}:{top[0][1]} {top[1][0][6:]}:{top[1][1]}")
