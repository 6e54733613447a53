"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / mean / total / share, and the
kernels of one PCG iteration in launch order.  Usage: python tools/launch_list_summary.py FILE.csv [--iteration]"""
import collections, csv, sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr, agg, seq = None, collections.defaultdict(list), []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            nm = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[-36:]
            v = float(d["Metric Value"].replace(",", ""))
            agg[nm].append(v); seq.append((nm, v))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:38s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.2f} us total={sum(v)/1e6:8.3f} ms {100*sum(v)/tot:5.1f}%")
print(f"total {tot/1e6:.3f} ms over {len(seq)} launches")
if "--iteration" in sys.argv:
    i0 = [i for i, (n, _) in enumerate(seq) if n.startswith("spmv")]
    if len(i0) > 3:
        a, b = i0[2], i0[3]
        print(f"one iteration: {sum(v for _, v in seq[a:b])/1e3:.1f} us, {b-a} launches")
        for n, v in seq[a:b]:
            print(f"   {n:38s} {v/1e3:7.2f}")
