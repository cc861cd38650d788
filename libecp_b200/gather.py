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


def allgather_shards(pack, unpack, rows, sizes, rank, world, device, group=None, buffers=None):
    """The collective step alone: `pack(rows, tensor)` fills the send buffer with this rank's packed rows,
    `unpack(rows, tensor)` scatters another rank's shard.  One all-gather of shards padded to the largest one.
    (The CUDA path passes the C-ABI kernels; the CPU tier runs the same code over gloo with numpy callables.)"""
    import torch
    import torch.distributed as dist

    cap = max(max(sizes), 1)  # all_gather wants equal shards: pad to the largest
    if buffers is None or buffers[0].numel() < cap:
        # torch.empty, not zeros: a fill kernel on torch's stream would not be ordered against the pack kernel, which
        # runs on the library's own non-blocking stream (unpack only reads the packed prefix of every shard)
        buffers = (torch.empty(cap, dtype=torch.float64, device=device),
                   torch.empty(world * cap, dtype=torch.float64, device=device))
    send, recv = buffers
    pack(rows[rank], send[:cap])
    dist.all_gather_into_tensor(recv[: world * cap], send[:cap], group=group)
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream().synchronize()
    for r in range(world):
        if r != rank:
            unpack(rows[r], recv[r * cap:(r + 1) * cap])
    return buffers


def allgather_matrix(h, rank, world, group=None, buffers=None):
    """Call after h.integrals_device() on every rank.  On return the handle's device matrix holds the full
    upper-triangular ECP matrix on every rank.  Returns (seconds, bytes received per rank, buffers); pass `buffers`
    back in to reuse the staging tensors."""
    import torch

    rows, sizes = shard_layout(h, world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    buffers = allgather_shards(lambda r, t: h.pack_rows(r, t.data_ptr(), t.numel()),
                               lambda r, t: h.unpack_rows(r, t.data_ptr(), t.numel()),
                               rows, sizes, rank, world, "cuda", group=group, buffers=buffers)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, 8 * (sum(sizes) - sizes[rank]), buffers


def numpy_pack(M, rows, out):
    """host mirror of k_pack_rows: the upper-triangle parts M[i][i:] of `rows`, back to back (tests, CPU tier)"""
    o = 0
    n = M.shape[0]
    for i in rows:
        out[o:o + n - i] = M[i, i:]
        o += n - i
    return o


def numpy_unpack(M, rows, packed):
    """host mirror of k_unpack_rows"""
    o = 0
    n = M.shape[0]
    for i in rows:
        M[i, i:] = packed[o:o + n - i]
        o += n - i
    return o
