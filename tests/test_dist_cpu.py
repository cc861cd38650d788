"""CPU tier: the N>1 path on world_size 2 with the gloo backend (no GPU): every rank builds the shard of
the executed-triple list it owns (host logic only - there is no CPU compute path) and the ranks check through
collectives that the shards are disjoint, complete and balanced, exactly what bench.py relies on when it lets each
GPU own a disjoint set of shell pairs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libecp_b200 import capi, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = synth.cfg3(4)
    with capi.Handle(s, tables_only=True) as h:
        full = h.triple_list()
        h.set_shard(rank, world)
        mine = h.triple_list()
    # key of a triple: (A,s1,B,s2,C) packed into one int64
    def key(t):
        t = t.astype(np.int64)
        return (((t[:, 0] * 64 + t[:, 1]) * 64 + t[:, 3]) * 64 + t[:, 4]) * 64 + t[:, 6]

    n_mine = torch.tensor([len(mine)], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, n_mine)
    total = int(sum(int(c) for c in counts))
    # gather the (padded) key lists and check disjointness + completeness on every rank
    pad = max(int(c) for c in counts)
    buf = torch.full((pad,), -1, dtype=torch.int64)
    buf[:len(mine)] = torch.from_numpy(key(mine))
    allk = [torch.empty(pad, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allk, buf)
    keys = np.concatenate([k.numpy()[:int(c)] for k, c in zip(allk, counts)])
    ok = (total == len(full)) and (len(np.unique(keys)) == len(full)) and set(keys.tolist()) == set(key(full).tolist())
    # every shell pair belongs to exactly one rank, for every centre
    pairs_mine = {(a, b, c, d) for a, b, _, c, d, _, _ in mine.tolist()}
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    share = torch.tensor([len(mine) / max(len(full), 1)], dtype=torch.float64)
    dist.all_reduce(share, op=dist.ReduceOp.MAX)
    # device-resident gather, host logic over gloo: every rank fills the rows it owns of a known matrix, one
    # all-gather of the padded packed shards, scatter of the other ranks' rows -> the full upper triangle everywhere
    from libecp_b200 import gather

    n = int(s["dim"])
    with capi.Handle(s, tables_only=True) as h:
        rows, sizes = gather.shard_layout(h, world)
        first = np.concatenate([[0], np.cumsum([(l + 1) * (l + 2) // 2 for l in s["lBS"]])])
        shell_first = np.concatenate([[0], np.cumsum(s["shellsBS"])])
        # AO rows of the row shells of my triples must be among the rows the layout gives me
        my_row_shells = np.unique(shell_first[mine[:, 0]] + mine[:, 1])
        mine_ok = all(set(range(first[x], first[x + 1])) <= set(rows[rank].tolist()) for x in my_row_shells)
    want = np.triu(np.arange(n * n, dtype=np.float64).reshape(n, n) + 0.5)
    M = np.zeros((n, n))
    for i in rows[rank]:
        M[i, i:] = want[i, i:]
    gather.allgather_shards(lambda r, t: gather.numpy_pack(M, r, t.numpy()), lambda r, t: gather.numpy_unpack(M, r, t.numpy()),
                            rows, sizes, rank, world, "cpu")
    gflag = torch.tensor([1 if (mine_ok and np.array_equal(M, want) and sum(sizes) == n * (n + 1) // 2) else 0],
                         dtype=torch.int64)
    dist.all_reduce(gflag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put((int(flag) * int(gflag), float(share), total, len(full), len(pairs_mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_shards_over_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    ok, share, total, nfull, _ = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == 1 and total == nfull
    assert share < 0.6  # the heavier rank holds less than 60 % of the triples
