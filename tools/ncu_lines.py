#!/usr/bin/env python
"""Warp-instructions per CUDA source line of one kernel in an .ncu-rep (ncu --set full; -lineinfo build; read here, no GPU).

    python tools/ncu_lines.py rep.ncu-rep <kernel substring> <divide by (e.g. rows)> [top N]
"""
import csv, io, subprocess, sys


def main():
    rep, pat, div = sys.argv[1], sys.argv[2], float(sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Function Name" and pat in r[1] and rows[i - 1][0] == "File Path" and rows[i - 1][1].endswith(".cu"):
            hdr = rows[i + 1]
            ie = hdr.index("Instructions Executed")
            per, txt = {}, {}
            j = i + 2
            while j < len(rows) and rows[j] and rows[j][0] not in ("Line No", "File Path", "Function Name"):
                q = rows[j]; j += 1
                if q[0] and len(q) > ie:
                    try:
                        per[int(q[0])] = per.get(int(q[0]), 0) + int(q[ie] or 0); txt[int(q[0])] = q[1].strip()[:110]
                    except ValueError:
                        pass
            print(f"# {r[1][:100]}: {sum(per.values()) / div:.1f} warp-instructions per unit")
            for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
                print(f"{v / div:8.1f} {k:5d} {txt[k]}")
            return
        i += 1
    print("kernel not found")


if __name__ == "__main__":
    main()
