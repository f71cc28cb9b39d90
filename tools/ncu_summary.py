"""Markdown summary of one `ncu --set full` capture, from its exported pages:
    ncu -i X.ncu-rep --page raw --csv > X_raw.csv
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_cs.csv     (optional)
usage: python tools/ncu_summary.py X_raw.csv [X_cs.csv] [--peak-tflops 33.7] [--title "..."]"""
import csv, sys, argparse
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_source_hotspots as hs

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores", "sm__icc_request_hit_rate.pct",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw"); ap.add_argument("cs", nargs="?")
    ap.add_argument("--peak-tflops", type=float, default=33.7)
    ap.add_argument("--title", default="")
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw, newline="")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {hdr[i]: (r[i], units[i]) for i in range(len(hdr))}
        name = d["Kernel Name"][0]
        print(f"## ncu --set full: `{name}`" + (f" — {a.title}" if a.title else "") + "\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k][0]} | {d[k][1]} |")
        stalls = []
        for k, (v, _) in d.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("\nWarp stall reasons (stalled warps per issued instruction): " +
              ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))
        try:
            cyc = float(d["smsp__cycles_elapsed.avg"][0])
            pc = lambda op: float(d[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"][0])
            flop = (pc("dadd") + pc("dmul") + 2.0 * pc("dfma")) * cyc
            ms = float(d["gpu__time_duration.sum"][0])
            if d["gpu__time_duration.sum"][1] in ("ns", "nsecond"): ms *= 1e-6
            elif d["gpu__time_duration.sum"][1] in ("us", "usecond"): ms *= 1e-3
            elif d["gpu__time_duration.sum"][1] in ("s", "second"): ms *= 1e3
            tf = flop / (ms * 1e-3) / 1e12
            print(f"\nExecuted FP64 flop (thread level, DADD + DMUL + 2 DFMA): {flop:.3e} in {ms:.1f} ms = **{tf:.2f} TFLOP/s** "
                  f"({100 * tf / a.peak_tflops:.1f} % of the measured {a.peak_tflops} TFLOP/s DFMA peak).  "
                  f"FP64 pipe busy {d['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'][0]} % of cycles at "
                  f"{d['smsp__thread_inst_executed_per_inst_executed.ratio'][0]} active threads per instruction.")
        except (KeyError, ValueError) as e:
            print(f"\n(executed-flop line unavailable: {e})")
        print()
    if a.cs:
        print("### Source hot spots (stall samples per CUDA line)\n")
        hs.main(a.cs, a.top)


if __name__ == "__main__":
    main()
