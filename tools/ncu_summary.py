"""Summarises .ncu-rep files (read with `ncu -i ... --page raw --csv`) into one line per captured launch."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]

def main(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0][-48:]
        vals = {}
        for k in KEYS:
            if k in idx:
                vals[k] = (r[idx[k]], units[idx[k]])
        def f(k, scale=1.0):
            if k not in vals: return float("nan")
            try: return float(vals[k][0].replace(",", "")) * scale
            except ValueError: return float("nan")
        dur = f("gpu__time_duration.sum"); u = vals.get("gpu__time_duration.sum", ("", ""))[1]
        dur_us = dur / 1e3 if u == "ns" else (dur if u == "us" else dur * 1e3 if u == "ms" else dur)
        def bytes_of(k):
            v = f(k); un = vals.get(k, ("", ""))[1].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(un, 1)
        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        print(f"{name:48s} {dur_us:9.1f} us  dram R {rd/1e6:8.1f} MB W {wr/1e6:8.1f} MB  -> {(rd+wr)/dur_us/1e3:7.0f} GB/s  "
              f"dram% {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}  sm% {f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}  "
              f"regs {f('launch__registers_per_thread'):4.0f}  warps% {f('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}  "
              f"L2 {bytes_of('lts__t_bytes.sum')/1e6:8.1f} MB  issue% {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}  "
              f"stall: lsb {f('smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct'):5.1f} lg {f('smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct'):5.1f} "
              f"bar {f('smsp__warp_issue_stalled_barrier_per_warp_active.pct'):5.1f} ssb {f('smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct'):5.1f} "
              f"mio {f('smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct'):5.1f}")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        main(p)
