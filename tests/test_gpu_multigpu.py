"""Needs two real GPUs (skipped otherwise): the drop-in CLI with `--gpus 2` — one worker thread per GPU, collectives over NCCL
inside the library (scs_nccl_init: ncclAllReduce for the per-pass counters and the weight vector, ncclAllGather for the genome /
amplicon replication), both workers writing ONE pair of files at their final offsets — must produce the files of `--gpus 1`."""
import os
import subprocess

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
EXE = os.path.join(H.ROOT, "scssim_b200", "bin", "scssim")


def _ngpu():
    from scssim_b200 import api
    return api.lib().scs_device_count()


@pytest.mark.parametrize("layout,isize,with_n", [("PE", 260, False), ("SE", 260, True), ("PE", 1200, False)])
def test_cli_two_real_gpus_over_nccl(tmp_path, layout, isize, with_n):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "cell.fa")
    if with_n:
        H.write_genome_with_n(fa, 300_000, 61)
    else:
        H.write_genome(fa, 2, 200_000, seed=61)          # 4 sequences -> 2 per worker
    prof = H.profile_path("Illumina_HiSeq2500")
    args = H.genreads_args(prof, layout, 3e-10, 10.0, isize)
    outs = {}
    for n in (1, 2):
        r = subprocess.run([EXE, "genreads", "-i", fa, "-o", os.path.join(tmp, f"g{n}"), "--seed", "99", "--gpus", str(n), "-t", "4"] + args,
                           capture_output=True, timeout=600, env=dict(os.environ, NCCL_DEBUG="WARN"))
        assert r.returncode == 0, r.stderr.decode()[-3000:]
        outs[n] = [H.read_bytes(p) for p in H.fastq_names(os.path.join(tmp, f"g{n}"), layout)]
    for a, b in zip(outs[1], outs[2]):
        assert len(a) > 100_000 and a == b


def test_cli_more_gpus_than_present_is_an_error(tmp_path):
    fa = os.path.join(str(tmp_path), "cell.fa")
    H.write_genome(fa, 1, 50_000, seed=1)
    prof = H.profile_path("Illumina_HiSeq2500")
    r = subprocess.run([EXE, "genreads", "-i", fa, "-m", prof, "-o", os.path.join(str(tmp_path), "x"), "--gpus", str(_ngpu() + 1)], capture_output=True, timeout=120)
    assert r.returncode != 0 and b"CUDA devices" in r.stderr


@pytest.mark.parametrize("gzip", [False, True])
def test_relay_through_a_peer_gpu_changes_nothing(tmp_path, gzip):
    """relay_device: the packed slabs travel over NVLink to GPU 1 and from there to the host (for boxes where some GPUs sit behind a
    slow host link); the bytes that land are those of the direct route — many slabs, callback sink and file sink."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    from scssim_b200 import api
    from scssim_b200.synth import synth_genome
    prof = H.profile_path("Illumina_HiSeq2500")
    genome = synth_genome(1, 300_000, seed=71, diploid=True)
    out = {}
    for relay in (-1, 1):
        with api.GenReads(gamma=3e-10, coverage=40.0, layout="PE", seed=5, slab_bytes=1 << 20, ring_slabs=3, gzip=gzip, device=0, relay_device=relay) as g:
            g.load_profile(prof).set_genome(genome).create_frags().amplify()
            a = g.yield_reads_bytes()
            g.yield_reads(os.path.join(str(tmp_path), f"r{relay}"))
            assert g.stats()["emit_launches"] >= 10
        ext = ".fq.gz" if gzip else ".fq"
        out[relay] = (a, [H.read_bytes(os.path.join(str(tmp_path), f"r{relay}_{k}{ext}")) for k in (1, 2)])
    assert out[1][0] == out[-1][0] and len(out[1][0][0]) > 100_000
    assert out[1][1] == out[-1][1] and out[1][1][0] == out[1][0][0]
