"""Multi-rank parity on one GPU: two ranks (threads, each with its own context and half of the cell's sequences,
collectives = in-process sums) must together write exactly the records a single rank writes for the whole cell —
same headers (global amplicon ids), same bases and qualities — because Philox streams and FASTQ headers are keyed by
global ids. Only the record order differs (each rank emits its own amplicons)."""
import os
import threading

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _records(fq: bytes):
    lines = fq.split(b"\n")
    assert lines[-1] == b""
    return sorted(b"\n".join(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4))


@pytest.mark.parametrize("layout,gamma,isize", [("PE", 2e-10, 260), ("SE", 5e-10, 260), ("PE", 2e-10, 1200)])
def test_two_ranks_equal_one_rank(tmp_path, layout, gamma, isize):
    from scssim_b200 import api
    from scssim_b200.dist import ThreadCollectives
    from scssim_b200.synth import synth_genome
    prof = H.profile_path("Illumina_HiSeq2500")
    genome = synth_genome(2, 150_000, seed=17, diploid=True)   # 4 sequences
    kw = dict(gamma=gamma, coverage=4.0, isize=isize, layout=layout, seed=777)
    with api.GenReads(**kw) as g:
        g.load_profile(prof).set_genome(genome).create_frags().amplify()
        one = g.yield_reads_bytes()
        st1 = g.stats()
    world = 2
    coll = ThreadCollectives(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            with api.GenReads(rank=rank, world=world, **kw) as g:
                g.set_collectives(*coll.pair())
                g.load_profile(prof).set_genome(genome[2 * rank:2 * rank + 2]).create_frags().amplify()
                out[rank] = (g.yield_reads_bytes(), g.stats())
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            coll.bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert sum(o[1]["n_fulls"] for o in out) == st1["n_fulls"]
    assert out[0][1]["n_fulls_global"] == st1["n_fulls"]
    for f in range(2 if layout == "PE" else 1):
        assert _records(out[0][0][f] + out[1][0][f]) == _records(one[f]), f"file {f}: shards differ from the single-rank output"
    assert len(one[0]) > 0


@pytest.mark.parametrize("layout,isize,weights", [("PE", 260, (1.0, 1.0)), ("PE", 1200, (1.0, 2.5)), ("SE", 260, (3.0, 1.0))])
def test_balanced_shards_concatenate_to_the_single_rank_files(tmp_path, layout, isize, weights):
    """balance=1: genome and amplicon table are replicated (device all-reduce hooks), the cell's read slots are cut by shard
    weight; rank-ordered concatenation of the shards must be byte-identical to the single-rank output."""
    from scssim_b200 import api
    from scssim_b200.dist import ThreadCollectives
    prof = H.profile_path("Illumina_HiSeq2500")
    genome = H.write_genome_with_n(os.path.join(str(tmp_path), "n.fa"), 200_000, 19)   # 2 haplotypes with N runs
    kw = dict(gamma=3e-10, coverage=5.0, isize=isize, layout=layout, seed=4321)
    with api.GenReads(**kw) as g:
        g.load_profile(prof).set_genome(genome).create_frags().amplify()
        one = g.yield_reads_bytes()
    world = 2
    coll = ThreadCollectives(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            with api.GenReads(rank=rank, world=world, balance=True, **kw) as g:
                g.set_collectives(*coll.pair())
                g.set_device_collective(*coll.device_pair())
                g.set_shard_weight(weights[rank])
                g.load_profile(prof).set_genome(genome[rank:rank + 1]).create_frags().amplify()
                out[rank] = (g.yield_reads_bytes(), g.stats())
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            coll.bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for f in range(2 if layout == "PE" else 1):
        assert out[0][0][f] + out[1][0][f] == one[f], f"file {f}"
    share = out[0][1]["records"] / (out[0][1]["records"] + out[1][1]["records"])
    assert abs(share - weights[0] / sum(weights)) < 0.02


@pytest.mark.parametrize("layout,isize,with_n", [("PE", 260, True), ("SE", 260, False), ("PE", 1200, False)])
def test_ranks_write_one_pair_of_files(tmp_path, layout, isize, with_n):
    """scs_yield_reads(prefix) with world = 2, balance = 1: each rank sizes its shard (scs_plan_fastq_bytes), the byte counts are
    exchanged, rank 0 creates the files and both ranks pwrite their shard at its final offset — the files must be byte-identical
    to the single-rank files (no shard files, no concatenation)."""
    from scssim_b200 import api
    from scssim_b200.dist import ThreadCollectives
    from scssim_b200.synth import synth_genome
    prof = H.profile_path("Illumina_HiSeq2500")
    tmp = str(tmp_path)
    genome = H.write_genome_with_n(os.path.join(tmp, "n.fa"), 200_000, 19) if with_n else synth_genome(1, 200_000, seed=29, diploid=True)
    kw = dict(gamma=3e-10, coverage=12.0, isize=isize, layout=layout, seed=99, slab_bytes=1 << 20, ring_slabs=3)
    with api.GenReads(**kw) as g:
        g.load_profile(prof).set_genome(genome).create_frags().amplify()
        g.yield_reads(os.path.join(tmp, "one"))
    world = 2
    coll = ThreadCollectives(world)
    errs = []

    def run(rank):
        try:
            with api.GenReads(rank=rank, world=world, balance=True, **kw) as g:
                g.set_collectives(*coll.pair())
                g.set_device_collective(*coll.device_pair())
                g.set_shard_weight((1.0, 1.7)[rank])
                g.load_profile(prof).set_genome(genome[rank:rank + 1]).create_frags().amplify()
                g.yield_reads(os.path.join(tmp, "two"))
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            coll.bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for a, b in zip(H.fastq_names(os.path.join(tmp, "two"), layout), H.fastq_names(os.path.join(tmp, "one"), layout)):
        assert os.path.getsize(b) > 100_000 and H.read_bytes(a) == H.read_bytes(b)
    assert not [f for f in os.listdir(tmp) if "rank" in f]
