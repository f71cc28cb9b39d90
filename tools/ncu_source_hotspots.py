"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.

usage: python tools/ncu_source_hotspots.py export.csv [top_n]
Prints, per source file and for the hottest lines, the stall samples, executed warp instructions and the
average number of active threads (lane efficiency)."""
import csv, sys, collections

def main(path, top=45):
    rows = list(csv.reader(open(path, newline="")))
    fname = None
    lines = []  # (file, line, src, samples, inst, thread_inst)
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if r[0] in ("Kernel Name", "Function Name") or hdr is None:
            continue
        if r[0] != "":
            try:
                lines.append((fname, int(r[0]), r[1].strip(), int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]),
                              int(r[hdr["Thread Instructions Executed"]]), int(r[hdr["stall_no_inst"]]), int(r[hdr["stall_wait"]]),
                              int(r[hdr["stall_long_sb"]])))
            except (ValueError, KeyError):
                pass
    tot_s = sum(l[3] for l in lines) or 1
    tot_i = sum(l[4] for l in lines) or 1
    tot_t = sum(l[5] for l in lines)
    print(f"total samples {tot_s}  warp instr {tot_i:.4g}  avg active threads {tot_t / tot_i:.2f}")
    byfile = collections.defaultdict(lambda: [0, 0, 0])
    for l in lines:
        b = byfile[l[0]]
        b[0] += l[3]; b[1] += l[4]; b[2] += l[5]
    print("\n| file | samples % | warp instr % | avg threads |\n|---|---:|---:|---:|")
    for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"| {f} | {100 * b[0] / tot_s:.1f} | {100 * b[1] / tot_i:.1f} | {b[2] / max(b[1], 1):.1f} |")
    print(f"\n| file:line | samples % | instr % | avg thr | no_inst % | wait % | long_sb % | source |\n|---|---:|---:|---:|---:|---:|---:|---|")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        s = max(l[3], 1)
        print(f"| {l[0]}:{l[1]} | {100 * l[3] / tot_s:.2f} | {100 * l[4] / tot_i:.2f} | {l[5] / max(l[4], 1):.1f} | "
              f"{100 * l[6] / s:.0f} | {100 * l[7] / s:.0f} | {100 * l[8] / s:.0f} | `{l[2][:90]}` |")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)


def by_ranges(path, fname, ranges):
    """ranges: list of (label, lo, hi) line ranges of one source file -> share of samples / instructions."""
    rows = list(csv.reader(open(path, newline="")))
    cur = None; hdr = None; tot_s = 0; tot_i = 0
    acc = collections.OrderedDict((lab, [0, 0, 0]) for lab, _, _ in ranges)
    acc["(other lines)"] = [0, 0, 0]
    for r in rows:
        if not r: continue
        if r[0] in ("File Name", "File Path"): cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
        if hdr is None or r[0] == "" or not r[0].isdigit(): continue
        try:
            s, i, t = int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]), int(r[hdr["Thread Instructions Executed"]])
        except ValueError:
            continue
        tot_s += s; tot_i += i
        if cur != fname: continue
        ln = int(r[0])
        for lab, lo, hi in ranges:
            if lo <= ln <= hi:
                a = acc[lab]; break
        else:
            a = acc["(other lines)"]
        a[0] += s; a[1] += i; a[2] += t
    print(f"\n| {fname} region | samples % | warp instr % | avg threads |\n|---|---:|---:|---:|")
    for lab, a in acc.items():
        print(f"| {lab} | {100 * a[0] / tot_s:.1f} | {100 * a[1] / tot_i:.1f} | {a[2] / max(a[1], 1):.1f} |")
