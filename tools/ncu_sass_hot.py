"""Where a kernel's warps wait: digest of an `ncu --page source --csv` export (SASS view).  Prints the stall-reason
totals, the instruction mix, and the hottest SASS instructions with their dominant stall.
Usage: python tools/ncu_sass_hot.py source.csv [top_n]"""
import csv
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data, tot, ex = [], 0, 0
    for k, r in enumerate(rows[hi + 1:]):
        try:
            s = int(r[ci["# Samples"]]); e = int(r[ci["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        data.append((s, e, k, r)); tot += s; ex += e
    print(f"instructions {len(data)}  samples {tot}  warp-instructions executed {ex}")
    agg = Counter()
    for s, e, k, r in data:
        for h in stalls:
            try:
                agg[h] += int(r[ci[h]])
            except ValueError:
                pass
    print("stall totals:", ", ".join(f"{h[6:]} {100 * v / max(tot, 1):.1f}%" for h, v in agg.most_common(9)))
    mix = Counter()
    for s, e, k, r in data:
        t = r[ci["Source"]].split()
        op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
        mix[op.split(".")[0]] += e
    print("executed mix:", ", ".join(f"{o} {100 * v / max(ex, 1):.1f}%" for o, v in mix.most_common(14)))
    print(f"--- top {top} by samples (index, samples, executed, dominant stall, sass)")
    for s, e, k, r in sorted(data, reverse=True)[:top]:
        dom = max(stalls, key=lambda h: int(r[ci[h]]) if r[ci[h]].isdigit() else 0)
        print(f"{k:5d} {s:6d} {e:9d} {dom[6:]:16s} {r[ci['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
