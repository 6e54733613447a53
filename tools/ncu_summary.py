"""Summarises .ncu-rep files (read with `ncu -i ... --page raw --csv`) into one line per captured launch: duration, DRAM
bytes and rate, SM / issue utilisation, registers, occupancy, predicated-lane efficiency and the warp-stall breakdown
(every `...issue_stalled_<reason>...` column the report carries, largest first).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [...]            # text, one line per launch
    python tools/ncu_summary.py --json out.json gpurun_out/prof.ncu-rep  # + {kernel: DRAM bytes per launch} for bench.py
"""
import csv
import io
import json
import re
import subprocess
import sys

STALL = re.compile(r"issue_stalled_([a-z_]+?)(?:_per_warp_active\.pct|_per_issue_active\.ratio|\.ratio|\.pct|$)")


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path, traffic):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    # one stall column per reason: prefer the "per issue active" ratio (cycles a warp waits per instruction it issues),
    # else the pc-sampling counts
    stall_cols = {}
    for h in hdr:
        if "issue_stalled" not in h or "not_issued" in h:
            continue
        m = STALL.search(h)
        if not m:
            continue
        reason = m.group(1)
        rank = 0 if h.startswith("smsp__average_warps_issue_stalled") and "per_issue_active" in h else (1 if "pcsamp" in h else 2)
        if reason not in stall_cols or rank < stall_cols[reason][0]:
            stall_cols[reason] = (rank, h)
    # one unit for all reasons: the per-issue-active ratios when the report has them (the pc-sampling columns are raw sample
    # counts and name some reasons in the plural, e.g. "no_instructions": mixing them in would dwarf every ratio)
    best = min((rk for rk, _ in stall_cols.values()), default=2)
    stall_cols = {k: v for k, v in stall_cols.items() if v[0] == best}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0][-48:]

        def f(k):
            return num(r[idx[k]]) if k in idx else float("nan")

        def bytes_of(k):
            if k not in idx:
                return float("nan")
            return f(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[idx[k]].lower(), 1)
        dur = f("gpu__time_duration.sum")
        u = units[idx["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in idx else ""
        dur_us = dur / 1e3 if u in ("ns", "nsecond") else (dur if u in ("us", "usecond") else dur * 1e3 if u in ("ms", "msecond") else dur)
        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        stalls = sorted(((num(r[idx[h]]), reason) for reason, (_, h) in stall_cols.items()), reverse=True)
        stalls = [(v, k) for v, k in stalls if v == v and v > 0][:6]
        tot = sum(v for v, _ in stalls) or 1.0
        stall_txt = " ".join(f"{k} {v:.2f}({100 * v / tot:.0f}%)" for v, k in stalls)
        print(f"{name:48s} {dur_us:9.1f} us  dram R {rd/1e6:8.1f} MB W {wr/1e6:8.1f} MB  -> {(rd+wr)/dur_us/1e3:7.0f} GB/s  "
              f"dram% {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}  sm% {f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}  "
              f"regs {f('launch__registers_per_thread'):4.0f}  warps% {f('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}  "
              f"issue% {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}  "
              f"lanes/inst {f('smsp__thread_inst_executed_per_inst_executed.ratio'):4.1f}  inst {f('smsp__inst_executed.sum')/1e6:8.1f} M  "
              f"stalls[{stall_txt}]")
        if traffic is not None and rd == rd:
            key = name.split("::")[-1].split("<")[0].strip()
            traffic.setdefault(key, []).append(rd + wr)


if __name__ == "__main__":
    args = sys.argv[1:]
    traffic, jpath = None, None
    if args and args[0] == "--json":
        jpath, args = args[1], args[2:]
        traffic = {}
    for p in args:
        print("==", p)
        main(p, traffic)
    if jpath:
        out = {k: sum(v) / len(v) for k, v in traffic.items()}
        out["source"] = "ncu --set full capture(s): " + ", ".join(args) + " (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)"
        json.dump(out, open(jpath, "w"), indent=1)
