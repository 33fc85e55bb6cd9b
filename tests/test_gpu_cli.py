"""The drop-in CLI (`scssim_b200/bin/scssim genreads ...`) end to end on the GPU: same flags as the reference, FASTQ files
byte-identical to the CPU oracle run with the same --seed; writes the .fai side file; SE and PE file naming."""
import os
import subprocess

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
EXE = os.path.join(H.ROOT, "scssim_b200", "bin", "scssim")


@pytest.mark.parametrize("layout", ["PE", "SE"])
def test_cli_matches_oracle(tmp_path, layout):
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "cell.fa")
    H.write_genome(fa, 1, 200_000, seed=31)
    prof = H.profile_path("Illumina_HiSeq2500")
    args = H.genreads_args(prof, layout, 2e-10, 5.0, 260)
    H.run_oracle(fa, os.path.join(tmp, "orc"), args, seed=4711)
    r = subprocess.run([EXE, "genreads", "-i", fa, "-t", "5", "-o", os.path.join(tmp, "gpu"), "--seed", "4711"] + args, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert b"MALBAC amplification..." in r.stderr and b"*****Producing reads*****" in r.stderr and b"Reads generation done!" in r.stderr
    assert os.path.exists(fa + ".fai")
    for a, b in zip(H.fastq_names(os.path.join(tmp, "gpu"), layout), H.fastq_names(os.path.join(tmp, "orc"), layout)):
        assert H.read_bytes(a) == H.read_bytes(b)


def test_cli_missing_profile_exits_like_reference(tmp_path):
    fa = os.path.join(str(tmp_path), "cell.fa")
    H.write_genome(fa, 1, 50_000, seed=1)
    r = subprocess.run([EXE, "genreads", "-i", fa, "-m", "/nonexistent.profile", "-o", os.path.join(str(tmp_path), "x")], capture_output=True)
    assert r.returncode != 0 and b"can not open file /nonexistent.profile" in r.stderr


def test_cli_two_gpu_workers_write_the_same_records(tmp_path):
    """--gpus 2 (both workers on this GPU via the test hook): the concatenated shards are byte-identical to --gpus 1."""
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "cell.fa")
    H.write_genome(fa, 2, 120_000, seed=33)          # 4 sequences -> 2 per worker
    prof = H.profile_path("Illumina_HiSeq2500")
    args = H.genreads_args(prof, "PE", 2e-10, 4.0, 260)
    outs = {}
    for n in (1, 2):
        r = subprocess.run([EXE, "genreads", "-i", fa, "-o", os.path.join(tmp, f"g{n}"), "--seed", "99", "--gpus", str(n)] + args,
                           capture_output=True, env=dict(os.environ, SCS_CLI_SAME_DEVICE="1"))
        assert r.returncode == 0, r.stderr.decode()
        outs[n] = [H.read_bytes(os.path.join(tmp, f"g{n}_{k}.fq")) for k in (1, 2)]
        assert not os.path.exists(os.path.join(tmp, f"g{n}.rank0_1.fq"))

    for k in range(2):   # the workers write contiguous ranges of the cell's read slots: identical files
        assert len(outs[1][k]) > 0 and outs[1][k] == outs[2][k]


def test_cli_more_workers_than_sequences(tmp_path):
    """A worker that owns no sequence still takes part in every collective; output equals the single-worker run."""
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "cell.fa")
    H.write_genome(fa, 1, 150_000, seed=35, diploid=False)      # one sequence, two workers
    prof = H.profile_path("Illumina_HiSeq2000")
    args = H.genreads_args(prof, "SE", 3e-10, 3.0, 260)
    outs = {}
    for n in (1, 2):
        r = subprocess.run([EXE, "genreads", "-i", fa, "-o", os.path.join(tmp, f"g{n}"), "--seed", "7", "--gpus", str(n)] + args,
                           capture_output=True, env=dict(os.environ, SCS_CLI_SAME_DEVICE="1"), timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        outs[n] = H.read_bytes(os.path.join(tmp, f"g{n}.fq"))
    assert len(outs[1]) > 0 and outs[1] == outs[2]


def test_fasta_layouts_load_identically(tmp_path):
    """Fixed-width FASTA (newlines skipped on the GPU), CRLF line ends, a single-line record and a ragged record (host
    gather) all pack to the same genome: identical FASTQ for the same seed."""
    from scssim_b200 import api
    from scssim_b200.synth import synth_genome
    tmp = str(tmp_path)
    genome = synth_genome(1, 90_000, seed=51, diploid=True)
    prof = H.profile_path("Illumina_HiSeq2500")

    def write(path, width, eol=b"\n", ragged=False):
        with open(path, "wb") as f:
            for name, s in genome:
                f.write(b">" + name.encode() + b" some description" + eol)
                b = s.tobytes()
                if ragged:
                    i, w = 0, 37
                    while i < len(b):
                        f.write(b[i:i + w] + eol); i += w; w = 37 + (i % 11)
                elif width is None:
                    f.write(b + eol)
                else:
                    for i in range(0, len(b), width):
                        f.write(b[i:i + width] + eol)
    outs = []
    for k, kw in enumerate([dict(width=60), dict(width=100, eol=b"\r\n"), dict(width=None), dict(width=None, ragged=True), dict(width=32)]):
        fa = os.path.join(tmp, f"g{k}.fa")
        write(fa, **{"width": kw.get("width"), "eol": kw.get("eol", b"\n"), "ragged": kw.get("ragged", False)})
        with api.GenReads(gamma=3e-10, coverage=3.0, layout="PE", seed=5) as g:
            g.load_profile(prof).load_genome(fa).create_frags().amplify()
            outs.append(g.yield_reads_bytes())
        fai = open(fa + ".fai").read().split("\n")[0].split("\t")
        assert fai[0] == genome[0][0] and int(fai[1]) == 90_000
    assert len(outs[0][0]) > 0
    for o in outs[1:]:
        assert o == outs[0]


def test_cli_with_every_buffer_on_virtual_memory(tmp_path):
    """SCS_BIG_ALLOC_BYTES=64 KiB pushes practically every device buffer onto the grow-in-place VMM path that human-scale runs
    use for their multi-GB arrays (vmm.h): same FASTQ as the oracle, simuvars included."""
    tmp = str(tmp_path)
    env = dict(os.environ, SCS_BIG_ALLOC_BYTES="65536")
    ref, snp, var = H.make_simuvars_case(tmp, 41, chrom_lens=(150_000, 60_000), n_cnv=3)
    H.run_oracle_simuvars(ref, snp, var, os.path.join(tmp, "orc_cell.fa"))
    r = subprocess.run([EXE, "simuvars", "-r", ref, "-s", snp, "-v", var, "-o", os.path.join(tmp, "cell.fa")], capture_output=True, env=env)
    assert r.returncode == 0, r.stderr.decode()
    assert H.read_bytes(os.path.join(tmp, "cell.fa")) == H.read_bytes(os.path.join(tmp, "orc_cell.fa"))
    prof = H.profile_path("Illumina_HiSeq2500")
    args = H.genreads_args(prof, "PE", 1e-9, 5.0, 260)          # default gamma: the lists grow over several passes
    fa = os.path.join(tmp, "cell.fa")
    H.run_oracle(fa, os.path.join(tmp, "orc"), args, seed=808)
    r = subprocess.run([EXE, "genreads", "-i", fa, "-o", os.path.join(tmp, "gpu"), "--seed", "808"] + args, capture_output=True, env=env)
    assert r.returncode == 0, r.stderr.decode()
    for a, b in zip(H.fastq_names(os.path.join(tmp, "gpu"), "PE"), H.fastq_names(os.path.join(tmp, "orc"), "PE")):
        assert len(H.read_bytes(a)) > 100_000 and H.read_bytes(a) == H.read_bytes(b)


def test_cli_gz_flag(tmp_path):
    """--gz: <prefix>_1.fq.gz / _2.fq.gz decompress (gzip -dc) to the oracle's files; with --gpus 2 one shard per worker whose
    concatenation in rank order decompresses to the same bytes."""
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "cell.fa")
    H.write_genome(fa, 2, 120_000, seed=33)
    prof = H.profile_path("Illumina_HiSeq2500")
    args = H.genreads_args(prof, "PE", 2e-10, 6.0, 260)
    H.run_oracle(fa, os.path.join(tmp, "orc"), args, seed=4711)
    want = [H.read_bytes(p) for p in H.fastq_names(os.path.join(tmp, "orc"), "PE")]
    r = subprocess.run([EXE, "genreads", "-i", fa, "-o", os.path.join(tmp, "g1"), "--seed", "4711", "--gz"] + args, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    for k, w in zip((1, 2), want):
        assert subprocess.run(["gzip", "-dc", os.path.join(tmp, f"g1_{k}.fq.gz")], capture_output=True, check=True).stdout == w
    r = subprocess.run([EXE, "genreads", "-i", fa, "-o", os.path.join(tmp, "g2"), "--seed", "4711", "--gz", "--gpus", "2"] + args, capture_output=True,
                       env=dict(os.environ, SCS_CLI_SAME_DEVICE="1"))
    assert r.returncode == 0, r.stderr.decode()
    for k, w in zip((1, 2), want):
        cat = b"".join(H.read_bytes(os.path.join(tmp, f"g2.rank{rk}_{k}.fq.gz")) for rk in (0, 1))
        assert subprocess.run(["gzip", "-dc"], input=cat, capture_output=True, check=True).stdout == w
