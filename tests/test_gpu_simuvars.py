"""simuvars on the GPU through the C ABI and the drop-in CLI: the FASTA is byte-identical to the CPU oracle's
(which is byte-identical to the reference binary's, tests/test_simuvars_cpu.py), for every FASTA layout; the device-resident
variant (`scs_simuvars_to_genome`) gives the genreads path exactly the genome the FASTA round trip would."""
import hashlib
import json
import lzma
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from scssim_b200 import api

pytestmark = pytest.mark.gpu
EXE = os.path.join(H.ROOT, "scssim_b200", "bin", "scssim")


@pytest.fixture(scope="module")
def ctx():
    with api.GenReads(device=0) as g:
        yield g


def test_golden_case(tmp_path, ctx):
    tmp = str(tmp_path)
    gold = os.path.join(H.GOLDEN, "simuvars_small")
    for f in ("ref.fa", "snp.txt", "vars.txt", "expected.fa"):
        with lzma.open(os.path.join(gold, f + ".xz")) as i, open(os.path.join(tmp, f), "wb") as o:
            o.write(i.read())
    ref, snp, var = (os.path.join(tmp, f) for f in ("ref.fa", "snp.txt", "vars.txt"))
    ctx.simuvars(ref, snp, var, os.path.join(tmp, "gpu.fa"))
    got = H.read_bytes(os.path.join(tmp, "gpu.fa"))
    assert hashlib.sha256(got).hexdigest() == json.load(open(os.path.join(gold, "meta.json")))["sha256"]
    assert got == H.read_bytes(os.path.join(tmp, "expected.fa"))
    st = ctx.simuvars_stats()
    assert st["launches"] >= 2 + 4 and st["out_bytes"] == len(got) and st["n_cnv"] == 10 and st["n_snp"] == 400
    assert os.path.exists(ref + ".fai")


@pytest.mark.parametrize("seed,width", [(1, 100), (2, 60), (3, 100), (11, 70), (12, 100), (13, 17)])
def test_matches_oracle(tmp_path, ctx, seed, width):
    d = str(tmp_path)
    ref, snp, var = H.make_simuvars_case(d, seed, width=width)
    H.run_oracle_simuvars(ref, snp, var, os.path.join(d, "orc.fa"))
    assert ctx.simuvars_bytes(ref, snp, var) == H.read_bytes(os.path.join(d, "orc.fa"))


def test_stress_cases_match_oracle_or_fail_like_it(tmp_path, ctx):
    from test_simuvars_cpu import _stress_params
    n_ok = 0
    for seed in range(100, 125):
        d = os.path.join(str(tmp_path), f"s{seed}")
        ref, snp, var = H.make_simuvars_case(d, seed, **_stress_params(seed))
        o = H.run_oracle_simuvars(ref, snp, var, os.path.join(d, "orc.fa"), check=False)
        if o.returncode != 0:
            with pytest.raises(api.ScsError):
                ctx.simuvars_bytes(ref, snp, var)
            continue
        assert ctx.simuvars_bytes(ref, snp, var) == H.read_bytes(os.path.join(d, "orc.fa")), seed
        n_ok += 1
    assert n_ok >= 12


def test_fasta_layouts_and_line_widths(tmp_path, ctx):
    """Reference FASTA as one line per record, CRLF-free ragged tail, and narrow lines; output width 100 regardless."""
    d = str(tmp_path)
    from scssim_b200.synth import synth_sequence, write_fasta
    seqs = [("chr1", synth_sequence(30_011, 1)), ("chr2", synth_sequence(16, 2)), ("chr3", synth_sequence(4_000, 3))]
    snp = os.path.join(d, "snp.txt")
    open(snp, "w").write("".join(f"rs{p}\tchr1\t{p}\tA/C\t+\tA\n" for p in range(5, 30_000, 37)) + "rs1\tchr2\t16\tA/G\t+\tA\n")
    var = os.path.join(d, "vars.txt")
    open(var, "w").write("i\tchr1\t100\tacgtn\thomo\nd\tchr1\t200\t150\thet\nc\tchr3\t1000\t2000\t4\t3\ni\tchr2\t16\ttt\thomo\nd\tchr2\t1\t3\thomo\n")
    outs = []
    for width in (100, 16, 1_000_000):
        ref = os.path.join(d, f"ref{width}.fa")
        write_fasta(ref, seqs, width)
        H.run_oracle_simuvars(ref, snp, var, os.path.join(d, f"orc{width}.fa"))
        got = ctx.simuvars_bytes(ref, snp, var)
        assert got == H.read_bytes(os.path.join(d, f"orc{width}.fa")), width
        outs.append(got)
    assert outs[0] == outs[1] == outs[2]


def test_large_chromosome_properties(tmp_path, ctx):
    """24 Mb chromosome, 20 k SNPs, 2 k indels, 12 CNVs: far beyond what the oracle's tests cover in seconds, so check
    size-independent properties — (1) byte equality with the oracle on this one case (the oracle still finishes in seconds
    at 24 Mb), (2) record lengths follow from the plan's copy/indel accounting, (3) running twice is idempotent,
    (4) without variants the haplotypes equal the upper-cased reference."""
    d = str(tmp_path)
    ref, snp, var = H.make_simuvars_case(d, 77, chrom_lens=(24_000_000,), n_snp=20_000, n_snv=200, n_ins=1000, n_del=1000, n_cnv=12)
    a = ctx.simuvars_bytes(ref, snp, var)
    st = ctx.simuvars_stats()
    b = ctx.simuvars_bytes(ref, snp, var)
    assert a == b
    H.run_oracle_simuvars(ref, snp, var, os.path.join(d, "orc.fa"))
    assert a == H.read_bytes(os.path.join(d, "orc.fa"))
    assert st["out_bytes"] == len(a) and st["out_bases"] + (st["out_bases"] + 99) // 100 <= len(a)
    plain = ctx.simuvars_bytes(ref, None, None)
    recs = H.read_fasta_records(ref)
    t = recs[0][1].tobytes().upper()
    body = b"".join(t[i:i + 100] + b"\n" for i in range(0, len(t), 100))
    assert plain == b">1_1_24000000\n" + body + b">1_2_24000000\n" + body


def test_to_genome_equals_fasta_round_trip(tmp_path):
    """simuvars -> device genome -> genreads  ==  simuvars -> FASTA file -> genreads (same seed): identical FASTQ."""
    d = str(tmp_path)
    ref, snp, var = H.make_simuvars_case(d, 21, chrom_lens=(150_000, 90_000), n_cnv=4)
    prof = H.profile_path("Illumina_HiSeq2500")
    with api.GenReads(gamma=2e-10, coverage=4.0, seed=77, device=0) as g:
        g.load_profile(prof)
        g.simuvars(ref, snp, var, os.path.join(d, "cell.fa"))
        g.load_genome(os.path.join(d, "cell.fa")).create_frags().amplify()
        via_file = g.yield_reads_bytes()
        n_file = g.stats()["genome_bases"]
    with api.GenReads(gamma=2e-10, coverage=4.0, seed=77, device=0) as g:
        g.load_profile(prof).simuvars_to_genome(ref, snp, var).create_frags().amplify()
        direct = g.yield_reads_bytes()
        assert g.stats()["genome_bases"] == n_file
    assert len(via_file[0]) > 100_000 and via_file == direct


def test_cli_simuvars_drop_in(tmp_path):
    d = str(tmp_path)
    ref, snp, var = H.make_simuvars_case(d, 31)
    H.run_oracle_simuvars(ref, snp, var, os.path.join(d, "orc.fa"))
    r = subprocess.run([EXE, "simuvars", "-r", ref, "-s", snp, "-v", var, "-o", os.path.join(d, "gpu.fa")], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert H.read_bytes(os.path.join(d, "gpu.fa")) == H.read_bytes(os.path.join(d, "orc.fa"))
    err = r.stderr.decode()
    assert "Details of the aberrations loaded from file" in err and "CNV: 12" in err and "600 SNPs to simulate were loaded" in err
    assert "Reference sequence was loaded from file" in err and "Elapsed time" in err
    # gz input is gunzipped beside the input like the reference does; missing -o is an argument error
    subprocess.run(["gzip", "-k", ref], check=True)
    os.remove(ref)
    r = subprocess.run([EXE, "simuvars", "--ref", ref + ".gz", "--var", var, "--output", os.path.join(d, "gpu2.fa")], capture_output=True)
    assert r.returncode == 0 and os.path.exists(ref), r.stderr.decode()
    assert b"Warning: SNP file not specified!" in r.stderr
    H.run_oracle_simuvars(ref, None, var, os.path.join(d, "orc2.fa"))
    assert H.read_bytes(os.path.join(d, "gpu2.fa")) == H.read_bytes(os.path.join(d, "orc2.fa"))
    r = subprocess.run([EXE, "simuvars", "-r", ref], capture_output=True)
    assert r.returncode == 1 and b"Use --output to specify the output file." in r.stderr
    r = subprocess.run([EXE, "simuvars", "-r", ref, "-v", os.path.join(d, "nope.txt"), "-o", os.path.join(d, "x.fa")], capture_output=True)
    assert r.returncode != 0 and b"can not open file" in r.stderr


def test_to_genome_two_ranks_equal_one_rank(tmp_path):
    """world = 2: every rank builds the same plan and materialises only its share of the haplotypes; together the ranks
    write exactly the records of the single-rank run (global ids key headers and Philox streams)."""
    import threading
    from scssim_b200.dist import ThreadCollectives
    d = str(tmp_path)
    ref, snp, var = H.make_simuvars_case(d, 23, chrom_lens=(120_000, 110_000), n_cnv=3)
    prof = H.profile_path("Illumina_HiSeq2500")
    kw = dict(gamma=2e-10, coverage=4.0, seed=55, layout="SE")
    with api.GenReads(**kw) as g:
        g.load_profile(prof).simuvars_to_genome(ref, snp, var).create_frags().amplify()
        one = g.yield_reads_bytes()[0]
        n_seq = g.stats()["n_sequences"]
    assert n_seq == 4
    world, coll, out, errs = 2, ThreadCollectives(2), [None, None], []

    def run(rank):
        try:
            with api.GenReads(rank=rank, world=world, **kw) as g:
                g.set_collectives(*coll.pair())
                g.load_profile(prof).simuvars_to_genome(ref, snp, var).create_frags().amplify()
                out[rank] = (g.yield_reads_bytes()[0], g.stats())
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            coll.bar.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert out[0][1]["n_sequences"] + out[1][1]["n_sequences"] == 4 and out[0][1]["n_sequences"] in (1, 2, 3)
    rec = lambda fq: sorted(b"\n".join(l) for l in zip(*[iter(fq.split(b"\n")[:-1])] * 4))
    assert len(one) > 50_000 and rec(out[0][0] + out[1][0]) == rec(one)


def test_config0_testdata_case(tmp_path, ctx):
    """BASELINE configs[0], simuvars half, on the GPU: the reference's own testData SNP / variation files on a synthetic chr20;
    the FASTA must hash to what the reference binary wrote (tests/golden/testdata/config0.json)."""
    d = str(tmp_path)
    meta = json.load(open(os.path.join(H.GOLDEN, "testdata", "config0.json")))
    ref, snp, var = H.config0_inputs(d)
    ctx.simuvars(ref, snp, var, os.path.join(d, "cell.fa"))
    data = H.read_bytes(os.path.join(d, "cell.fa"))
    assert len(data) == meta["bytes"] and hashlib.sha256(data).hexdigest() == meta["sha256"]
    st = ctx.simuvars_stats()
    assert st["n_snp"] == 38603 and (st["n_cnv"], st["n_snv"], st["n_ins"], st["n_del"]) == (6, 11, 6, 6)
    recs = H.read_fasta_records(os.path.join(d, "cell.fa"))
    assert [(n, len(s)) for n, s in recs] == [tuple(r) for r in meta["records"]]
