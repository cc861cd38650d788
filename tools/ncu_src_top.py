"""Top-N source lines by warp-stall samples from `ncu --page source --csv` (needs -lineinfo)."""
import csv
import sys


def main(path, n):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if not rows:
        print("empty")
        return
    cols = [c for c in rows[0].keys() if c]
    samp = next((c for c in cols if "Sampling" in c and "All" in c), None) or next((c for c in cols if "Samples" in c), None)
    src = next((c for c in cols if c.strip() in ("Source", "source")), None)
    inst = next((c for c in cols if "Instructions Executed" in c), None)
    print("columns:", [c for c in cols][:40])

    def num(x):
        try:
            return float(str(x).replace(",", ""))
        except Exception:
            return 0.0

    tot = sum(num(r.get(samp)) for r in rows) or 1.0
    rows.sort(key=lambda r: -num(r.get(samp)))
    print(f"total samples {tot:.0f}")
    for r in rows[:n]:
        stall_cols = [(c, num(r[c])) for c in cols if c.startswith("stall_") or "Stall" in c]
        stall_cols = sorted([x for x in stall_cols if x[1] > 0], key=lambda x: -x[1])[:3]
        print(f"{num(r.get(samp)) / tot:6.1%} inst={r.get(inst, '')!s:>10} | {str(r.get(src, ''))[:110]} | {stall_cols}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
