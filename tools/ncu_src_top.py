"""Top-N SASS instructions by warp-stall samples from `ncu --page source --csv` (one kernel per file)."""
import csv
import gzip
import sys


def main(path, n):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt", newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    if lines and lines[0].startswith('"Kernel Name"'):
        print(lines[0].strip()[:150])
        lines = lines[1:]
    rows = list(csv.DictReader(lines))
    if not rows:
        print("empty")
        return

    def num(x):
        try:
            return float(str(x).replace(",", ""))
        except Exception:
            return 0.0

    samp = "Warp Stall Sampling (All Samples)"
    tot = sum(num(r.get(samp)) for r in rows) or 1.0
    inst = sum(num(r.get("Instructions Executed")) for r in rows)
    print(f"total samples {tot:.0f}; SASS lines {len(rows)}; warp instructions executed {inst:.0f}")
    stall_cols = [c for c in rows[0].keys() if c and c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(num(r[c]) for r in rows) for c in stall_cols}
    print("stall totals:", ", ".join(f"{c[6:]} {v / tot:.1%}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    # instruction mix
    mix = {}
    for r in rows:
        opn = str(r.get("Source", "")).strip().split()[0:2]
        opn = opn[1] if opn and opn[0].startswith("@") and len(opn) > 1 else (opn[0] if opn else "?")
        opn = opn.split(".")[0]
        mix[opn] = mix.get(opn, 0.0) + num(r.get("Instructions Executed"))
    print("executed mix:", ", ".join(f"{k} {v / max(inst, 1):.1%}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14]))
    order = sorted(range(len(rows)), key=lambda i: -num(rows[i].get(samp)))
    for i in order[:n]:
        r = rows[i]
        top = sorted([(c[6:], num(r[c])) for c in stall_cols if num(r[c]) > 0], key=lambda x: -x[1])[:3]
        print(f"{num(r.get(samp)) / tot:6.1%} line {i:5d} exec={num(r.get('Instructions Executed')):>10.0f} thr={num(r.get('Avg. Threads Executed')):4.1f} | "
              f"{str(r.get('Source', '')).strip()[:70]:70s} | {top}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
