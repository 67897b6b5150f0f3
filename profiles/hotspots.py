#!/usr/bin/env python
"""Per-instruction view of an ncu report captured with --import-source on (SASS page): executed warp-instructions and
warp-stall samples per instruction, in program order, with the hottest ones marked.

usage: python profiles/hotspots.py <report.ncu-rep> <kernel-regex> [min_share]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    ia, isrc, isamp, iex, ithr = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
    body = []
    for r in rows[hdr + 1:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            if body:
                break  # next kernel instance
            continue
        body.append((r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), r[ithr]))
    tot_s, tot_e = sum(b[1] for b in body), sum(b[2] for b in body)
    print(f"# {kern}: {len(body)} SASS instructions, {tot_e} warp-instructions executed, {tot_s} stall samples")
    acc_e = 0
    for i, (src, s, e, thr) in enumerate(body):
        acc_e += e
        mark = "*" if s >= min_share * tot_s else " "
        if mark == "*" or "--all" in sys.argv:
            print(f"{i:5d} {mark} samp {s:6d} ({100.0 * s / max(tot_s, 1):5.2f}%) exec {e:9d} thr {thr:>5s} cumexec {100.0 * acc_e / max(tot_e, 1):5.1f}%  {src}")


if __name__ == "__main__":
    main()
