"""Algorithmic-flop counts per config from the instrumented oracle restatement (oracle/oracle_ecp.c),
committed as tests/golden/falg.json and used by bench.py for roofline.achieved.
Run in the build container:  python tools/make_falg.py   (about 3 min)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libecp_b200 import synth  # noqa: E402
from oracle.refbind import algorithmic_flops, port_counters  # noqa: E402

CFG5_SAMPLE = [0, 1, 62, 63, 124, 125, 187, 188, 250, 251, 288, 289, 312, 313, 375, 376, 437, 438, 498, 499]


def one(name, s, out):
    t = time.time()
    rc, M, c = port_counters(s)
    k = algorithmic_flops(c)
    out[name] = dict(rc=rc, nominal=synth.nominal_triples(s), executed=c["triples_exec"], counters=c, flops=k,
                     oracle_seconds=time.time() - t, sum=float(M.sum()))
    print(name, "exec", c["triples_exec"], "flops/exec-triple %.1f k" % (k["total"] / max(c["triples_exec"], 1) / 1e3),
          {a: round(b / k["total"], 3) for a, b in k.items()}, "%.1fs" % (time.time() - t), flush=True)


def main():
    out = {}
    one("cfg1", synth.cfg1(), out)
    one("cfg2", synth.cfg2(), out)
    one("au2", synth.cfg3(2), out)
    one("au4", synth.cfg3(4), out)
    one("cfg4a", synth.cfg4("a"), out)
    one("cfg4b", synth.cfg4("b"), out)
    for c in CFG5_SAMPLE:
        one(f"cfg5_c{c}", synth.cfg5(500, active=[c]), out)
    one("cfg3", synth.cfg3(20), out)
    with open(os.path.join(ROOT, "tests", "golden", "falg.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
