"""world_size-2 gloo test (CPU) of the multi-rank host logic: the all-reduce adapters handed to
scs_set_collectives and the list-geometry arithmetic that turns per-rank product counts into global amplicon indices."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_index(lend, ltot, gbase, gtot, before, t):
    """Python restatement of scs::global_index (scssim_b200/csrc/ctx.h)."""
    b = 0
    while b + 1 < len(lend) and t >= lend[b]:
        b += 1
    q = t - (lend[b - 1] if b else 0)
    crank = before[b] + (ltot[b] - 1 - q)
    return gbase[b] + (gtot[b] - 1 - crank)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scssim_b200.dist import make_collectives
    ar_u64, ar_f64 = make_collectives(dist, device="cpu")
    # 1. sums
    a = np.array([rank + 1, 2 ** 63 + 5 * rank, 0], dtype=np.uint64)
    ar_u64(a)
    f = np.array([0.5 * (rank + 1), 1e-300], dtype=np.float64)
    ar_f64(f)
    # 2. the per-batch geometry exchange of amplify(): every rank contributes its product count per batch
    local = [[3, 0, 5], [2, 4, 1]][rank]                  # products of this rank in batches 0..2
    per_rank = np.zeros((3, world), dtype=np.uint64)
    per_rank[:, rank] = local
    flat = per_rank.reshape(-1).copy()
    ar_u64(flat)
    per_rank = flat.reshape(3, world)
    gtot = per_rank.sum(axis=1).tolist()
    before = per_rank[:, :rank].sum(axis=1).tolist()
    lend = np.cumsum(local).tolist()
    gbase = [0] + np.cumsum(gtot).tolist()[:-1]
    gidx = [_global_index(lend, local, gbase, gtot, before, t) for t in range(sum(local))]
    q.put((rank, a.tolist(), f.tolist(), gidx, int(sum(gtot))))
    dist.destroy_process_group()


def test_collectives_and_list_geometry_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, a, f, gidx, total in res:
        assert a == [3, (2 ** 63 + 2 ** 63 + 5) % 2 ** 64, 0]
        assert f[0] == 1.5
    # the two ranks' global indices partition [0, total) and, inside a batch, higher ranks come first (reverse creation order)
    all_idx = sorted(res[0][3] + res[1][3])
    assert all_idx == list(range(res[0][4]))
    assert res[0][3][:3] == [2 + 0, 2 + 1, 2 + 2] and res[1][3][:2] == [0, 1]   # batch 0: rank 1's 2 products precede rank 0's 3


def _shard_worker(rank, world, port, q):
    """Every rank works out its own share of the cell's haplotypes (as scs_simuvars_to_genome / scs_load_genome do) and
    the shares are gathered over gloo: together they must cover every sequence exactly once, in rank order."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scssim_b200 import api
    lens = [248_956_422, 248_956_422, 242_193_529, 242_193_529, 198_295_559, 198_295_559, 1000, 0, 57_227_415]
    lo, hi = api.shard_sequences(lens, rank, world)
    mine = torch.tensor([lo, hi, sum(lens[lo:hi])], dtype=torch.int64)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    q.put((rank, [o.tolist() for o in out], len(lens), sum(lens)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sequence_sharding_partitions_the_cell(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    _, shares, n, total = res[0]
    assert all(r[1] == shares for r in res)                      # every rank sees the same picture
    assert shares[0][0] == 0 and shares[-1][1] == n
    for a, b in zip(shares, shares[1:]):
        assert a[1] == b[0]                                       # contiguous, in rank order, no gaps or overlaps
    assert sum(s[2] for s in shares) == total
    assert max(s[2] for s in shares) < 1.6 * total / world        # about balanced by bases


def test_sequence_sharding_more_ranks_than_sequences():
    from scssim_b200 import api
    got = [api.shard_sequences([100, 100], r, 5) for r in range(5)]
    owned = [g for g in got if g[1] > g[0]]
    assert sorted(owned) == [(0, 1), (1, 2)] and all(g == (0, 0) for g in got if g[1] == g[0])


def _file_worker(rank, world, port, q, path):
    """The host side of scs_yield_reads with world > 1, on CPU: exchange shard sizes (all-reduce of a zero-padded vector), rank 0
    creates and preallocates the file, a second all-reduce is the barrier, every rank writes its shard at its offset through the
    asynchronous sink."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scssim_b200 import api
    from scssim_b200.dist import make_collectives
    ar_u64, _ = make_collectives(dist, device="cpu")
    rng = np.random.default_rng(100 + rank)
    mine = rng.integers(0, 256, size=[700_001, 13, 1_250_000][rank], dtype=np.uint8).tobytes()
    v = np.zeros(world, dtype=np.uint64); v[rank] = len(mine)
    ar_u64(v)
    off, total = int(v[:rank].sum()), int(v.sum())
    if rank == 0:
        api.write_file_async(path, mine, 64_000, base=0, create=True, prealloc=total)
    b = np.zeros(1, dtype=np.uint64); ar_u64(b)      # the file exists before the others open it
    if rank != 0:
        api.write_file_async(path, mine, 64_000, base=off, create=False)
    b = np.zeros(1, dtype=np.uint64); ar_u64(b)
    q.put((rank, mine))
    dist.destroy_process_group()


def test_ranks_write_one_file_at_exchanged_offsets(tmp_path):
    world, port = 3, _free_port()
    path = os.path.join(str(tmp_path), "shared.fq")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_file_worker, args=(r, world, port, q, path)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert open(path, "rb").read() == b"".join(m for _, m in res)
