"""Lane / chunk statistics of k_type1S from a diagnostic build (tools/build_variant.sh t1stats -DT1_STATS):
LIBECP_B200_SO=libecp_b200/lib/libecp_b200_t1stats.so python tools/t1_stats.py [workload ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libecp_b200 import capi, synth  # noqa: E402

NAMES = ["warp_iters", "group_chunks", "gc_c0", "gc_c1", "gc_levels", "lanes_inwin", "lanes_live", "live_c0", "live_c1",
         "live_levels", "pairs", "pairs_failed", "chunks_of_failed", "chunks_of_converged"]


def main():
    for wl in sys.argv[1:] or ["cfg3", "cfg5_60"]:
        s = synth.cfg3(20) if wl == "cfg3" else synth.cfg5(int(wl.split("_")[1]))
        buf = (C.c_ulonglong * 16)()
        with capi.Handle(s) as h:
            h.integrals_device()
            capi.lib().ecpdev_t1stats(buf)
            h.integrals_device()
            rc = capi.lib().ecpdev_t1stats(buf)
        v = dict(zip(NAMES, list(buf)))
        print(wl, "rc", rc, v)
        if rc == 1 and v["warp_iters"]:
            gc = v["group_chunks"]
            print(f"  groups with a pair per warp iteration {gc / v['warp_iters']:.2f} of 4")
            print(f"  chunks: c0 {v['gc_c0'] / gc:.1%}, c1 {v['gc_c1'] / gc:.1%}, levels>=4 {v['gc_levels'] / gc:.1%}")
            print(f"  live lanes per group-chunk: c0 {v['live_c0'] / max(v['gc_c0'], 1):.2f}, c1 {v['live_c1'] / max(v['gc_c1'], 1):.2f}, "
                  f"levels {v['live_levels'] / max(v['gc_levels'], 1):.2f} of 8; in-window {v['lanes_inwin'] / gc:.2f}; live per warp iteration {v['lanes_live'] / v['warp_iters']:.1f} of 32")
            print(f"  pairs {v['pairs']}, failed on the small grid {v['pairs_failed'] / max(v['pairs'], 1):.1%}; chunks per failed pair "
                  f"{v['chunks_of_failed'] / max(v['pairs_failed'], 1):.1f}, per converged pair {v['chunks_of_converged'] / max(v['pairs'] - v['pairs_failed'], 1):.1f}; "
                  f"share of chunks spent on failed pairs {v['chunks_of_failed'] / gc:.1%}")


if __name__ == "__main__":
    main()
