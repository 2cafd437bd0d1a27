"""Per-source-line instruction and stall shares of one kernel from an ncu report captured with --import-source on.
usage: source_hotspots.py <report.ncu-rep> <kernel name> [top N]   (prints markdown)"""
import csv
import io
import subprocess
import sys

STALLS = ["stall_wait", "stall_selected", "stall_branch_resolving", "stall_short_sb", "stall_long_sb", "stall_lg", "stall_mio",
          "stall_math", "stall_barrier", "stall_no_inst", "stall_dispatch", "stall_not_selected"]


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    cur, hdr, lines, seen = None, None, [], set()
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
            d = dict(zip(hdr, r))
            key = (cur, r[0])
            if key in seen:                      # a second launch of the same kernel repeats the listing
                continue
            seen.add(key)
            ex = int(d.get("Instructions Executed") or 0)
            if ex:
                lines.append((ex, int(d.get("Warp Stall Sampling (All Samples)") or 0), cur, r[0], r[1].strip(),
                              {k: int(d.get(k) or 0) for k in STALLS}))
    ti, ts = sum(l[0] for l in lines), sum(l[1] for l in lines)
    agg = {k: sum(l[5][k] for l in lines) for k in STALLS}
    print(f"### `{kern}`\n")
    print(f"warp instructions {ti / 1e6:.1f} M, stall samples {ts}; by reason: " +
          ", ".join(f"{k[6:]} {100 * v / max(ts, 1):.0f} %" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 50 > ts) + "\n")
    print("| inst % | stall % | line | source |\n|---|---|---|---|")
    for l in sorted(lines, reverse=True)[:top]:
        print(f"| {100 * l[0] / ti:.1f} | {100 * l[1] / max(ts, 1):.1f} | {l[2]}:{l[3]} | `{l[4][:110].replace('|', '¦')}` |")
    print()


if __name__ == "__main__":
    main()
