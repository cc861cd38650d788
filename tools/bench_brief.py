"""Print the key figures of bench.py JSON lines read from stdin (one per line)."""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("reference: value %.3e" % d["value"])
        continue
    msg = "value %.3e (%.1f ms/step) e2e %.3e (%.1f ms) build %.1f device %.1f launches %d" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["host_ms_build_per_step"],
        d["device_ms_per_step"], d["gpu_launches"])
    if "roofline" in d:
        r = d["roofline"]
        msg += "\n  serial %.1f %s\n  dominant %s achieved %.3f TF frac %.4f whole-step frac %.4f" % (
            r["device_ms_serial_step"], r["kernel_ms_per_step"], r["kernel"], r["achieved"], r["frac"], r["whole_step_frac"])
    if "secondary" in d:
        s = d["secondary"]
        msg += "\n  Au20 %.2f ms/step device %.2f %s whole-step frac %.4f" % (
            s["ms_per_step"], s["device_ms_per_step"], s["roofline"]["kernel_ms_per_step"], s["roofline"]["whole_step_frac"])
    if d.get("parity"):
        p = d["parity"]
        msg += "\n  parity ok=%s max_abs %.2e max_rel %.2e %s" % (p.get("ok"), p.get("max_abs", -1), p.get("max_rel", -1), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in p.items() if k.endswith("_worst") or k in ("block_support_equal", "sample_violations")})
    if d.get("allgather"):
        msg += "\n  allgather %.2f ms %.0f GB/s per rank; gathered step %.2f ms value %.3e" % (
            d["allgather"]["ms"], d["allgather"]["GBps_per_rank"], d["gathered"]["ms_per_step"], d["gathered"]["value"])
    if d.get("per_rank"):
        msg += "\n  per_rank %s" % d["per_rank"]
    print(msg)
