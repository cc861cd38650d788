"""Device-resident gather of a sharded ECP matrix.

Every rank of a sharded run holds, in its handle's device matrix, the upper-triangle parts of the AO rows of the
shells it owns (disjoint between ranks, no data-path collective during the integration).  A consumer that wants the
whole matrix on every GPU - the device-side counterpart of the reference's host callback accumulating into one matrix
(reference src/getIntegrals.c:22-43) - packs its rows, runs ONE all-gather over NCCL (NVLink / NVSwitch) and scatters
the other ranks' rows into its own matrix.  The pack / scatter kernels sit behind the C ABI
(libecp_b200_pack_rows / libecp_b200_unpack_rows, include/libecp_b200.h); the communicator is the caller's, here
torch.distributed.
"""
from __future__ import annotations

import time


def shard_layout(h, world):
    """(rows per rank, packed doubles per rank): a pure function of the basis and `world`, identical on every rank"""
    rows = [h.owned_rows(r, world) for r in range(world)]
    return rows, [h.packed_size(x) for x in rows]


def allgather_matrix(h, rank, world, group=None, buffers=None):
    """Call after h.integrals_device() on every rank.  On return the handle's device matrix holds the full
    upper-triangular ECP matrix on every rank.  Returns (seconds, bytes received per rank, buffers); pass `buffers`
    back in to reuse the staging tensors."""
    import torch
    import torch.distributed as dist

    rows, sizes = shard_layout(h, world)
    cap = max(sizes)  # all_gather wants equal shards: pad to the largest
    if buffers is None or buffers[0].numel() < cap:
        buffers = (torch.zeros(cap, dtype=torch.float64, device="cuda"),
                   torch.empty(world * cap, dtype=torch.float64, device="cuda"))
    send, recv = buffers
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h.pack_rows(rows[rank], send.data_ptr(), cap)
    dist.all_gather_into_tensor(recv[: world * cap], send[:cap], group=group)
    torch.cuda.current_stream().synchronize()
    for r in range(world):
        if r != rank:
            h.unpack_rows(rows[r], recv[r * cap:].data_ptr(), cap)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, 8 * (sum(sizes) - sizes[rank]), buffers
