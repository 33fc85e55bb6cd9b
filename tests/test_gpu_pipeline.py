"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Free-running mode: oracle and CUDA path draw from the same Philox4x32-10 streams, so every
intermediate array and the FASTQ must be identical byte for byte.
"""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _run_case(tmp, name, n_chrom, chrom_len, gseed, profile, layout, gamma, coverage, isize, seed, primers=100000, diploid=True, with_n=False,
              slab_bytes=0, min_slabs=0):
    from scssim_b200 import api
    fa = os.path.join(tmp, f"{name}.fa")
    genome = H.write_genome_with_n(fa, chrom_len, gseed) if with_n else H.write_genome(fa, n_chrom, chrom_len, gseed, diploid=diploid)
    prof = profile if os.path.isabs(profile) else H.profile_path(profile)
    args = H.genreads_args(prof, layout, gamma, coverage, isize, primers)
    oprefix, dprefix = os.path.join(tmp, name + "_orc"), os.path.join(tmp, name + "_dump")
    H.run_oracle(fa, oprefix, args, seed=seed, dump_prefix=dprefix)
    od = H.oracle_dump(dprefix)
    with api.GenReads(primers=primers, gamma=gamma, coverage=coverage, isize=isize, layout=layout, seed=seed, slab_bytes=slab_bytes) as g:
        g.load_profile(prof).load_genome(fa).create_frags()
        fr = g.dump(api.DUMP_FRAGS)
        assert np.array_equal(fr[:, :4], od["frags"][:, :4]), "fragments differ"
        g.amplify()
        st = g.stats()
        assert st["n_semis"] == len(od["semis"]) and st["n_fulls"] == len(od["fulls"]), (st, len(od["semis"]), len(od["fulls"]))
        esemis, efulls = H.expected_windows(genome, od)
        gs, gf = g.dump(api.DUMP_SEMIS).astype(np.int64), g.dump(api.DUMP_FULLS).astype(np.int64)
        assert np.array_equal(gs[:, :4], esemis[:, :4]), "semi amplicons differ"
        assert np.array_equal(gs[:, 4], esemis[:, 4]), "semi primer counts differ"
        assert np.array_equal(gf[:, :4], efulls[:, :4]), "full amplicons differ"
        assert np.array_equal(g.dump(api.DUMP_PRIMER_COUNTS), od["primer_counts"]), "primer pool differs"
        g.set_read_counts()
        assert np.allclose(g.dump(api.DUMP_WEIGHTS), od["weights"], rtol=1e-12, atol=0), "weights differ"
        assert np.array_equal(g.dump(api.DUMP_COUNTS), od["counts"]), "read counts differ"
        f1, f2 = g.yield_reads_bytes()
        st = g.stats()
    names = H.fastq_names(oprefix, layout)
    assert f1 == H.read_bytes(names[0]), "FASTQ file 1 differs"
    if layout == "PE":
        assert f2 == H.read_bytes(names[1]), "FASTQ file 2 differs"
    assert st["emit_launches"] >= min_slabs, f"expected at least {min_slabs} slabs, ran {st['emit_launches']}"
    return st


def test_pe_hiseq2500_small(tmp_path):
    st = _run_case(str(tmp_path), "pe2500", 1, 400_000, 7, "Illumina_HiSeq2500", "PE", 2e-10, 5.0, 260, seed=0x5C55)
    assert st["fastq_bytes"][0] > 0


def test_se_hiseq2000_default_gamma(tmp_path):
    _run_case(str(tmp_path), "se2000", 2, 150_000, 3, "Illumina_HiSeq2000", "SE", 1e-9, 3.0, 260, seed=12345)


def test_pe_xten_insert_failures(tmp_path):
    # -s 1200: insert sizes often exceed the amplicon -> fragCount gaps and the >1000-failures rule
    _run_case(str(tmp_path), "xten", 1, 400_000, 7, "Illumina_HiSeqXTen", "PE", 2e-10, 4.0, 1200, seed=99)


def test_pe_gaiix_short_insert(tmp_path):
    _run_case(str(tmp_path), "gaiix", 1, 400_000, 9, "Illumina_GenomeAnalyzerIIx", "PE", 5e-10, 4.0, 100, seed=4242)


def test_pe_mostly_failing_inserts(tmp_path):
    _run_case(str(tmp_path), "fail", 1, 400_000, 7, "Illumina_HiSeq2500", "PE", 3e-10, 6.0, 1900, seed=5)


def test_se_many_primers_per_fragment(tmp_path):
    # gamma 2.5e-9 on two 40 kb fragments: hundreds of primers per fragment (attached-site bitmap path instead of the
    # short list), ~70 k full amplicons from 3 k semi amplicons, nearly all reads allocated by the multinomial remainder
    st = _run_case(str(tmp_path), "manyprimers", 1, 40_000, 41, "Illumina_HiSeq2000", "SE", 2.5e-9, 2.0, 260, seed=5, diploid=False)
    assert st["n_fulls"] > 50_000


def test_se_more_primers_than_a_lane_holds(tmp_path):
    # gamma 5e-9: ~30 primers per semi amplicon, so most templates of the semi-amplicon passes exceed the 24 attached sites a lane
    # of amplify_semis_lanes_kernel keeps and are amplified by the warp kernel in the same pass; the rest by the lane kernel
    st = _run_case(str(tmp_path), "lanecap", 1, 30_000, 43, "Illumina_HiSeq2000", "SE", 5e-9, 1.0, 260, seed=6, diploid=False)
    assert st["n_fulls"] > 20_000


@pytest.mark.parametrize("layout", ["PE", "SE"])
def test_genome_with_n_iupac_and_lowercase(tmp_path, layout):
    # N runs / scattered N / IUPAC codes (all "N" after the reference's complement) / soft-masked bases: primer sites over N
    # never bind, countGC() = 0 for windows with N, reads over N emit 'N' with Q in [33,53) and consume draws differently
    st = _run_case(str(tmp_path), "withn" + layout, 1, 300_000, 77, "Illumina_HiSeq2500", layout, 3e-10, 6.0, 260, seed=9, with_n=True)
    assert st["records"] > 0


def test_degenerate_inputs_match_the_oracle(tmp_path):
    """Edge cases: (1) only sequences shorter than the 1027-base amplification minimum -> no amplicons, empty FASTQ (the
    reference itself crashes on an empty amplicon list; oracle and library write empty files); (2) a coverage that rounds to
    zero reads; (3) an empty record between two sequences and a last line without a terminator."""
    from scssim_b200 import api
    from scssim_b200.synth import synth_sequence, write_fasta
    d = str(tmp_path)
    prof = H.profile_path("Illumina_HiSeq2500")

    def both(fa, tag, layout, gamma, cov):
        args = H.genreads_args(prof, layout, gamma, cov, 260)
        H.run_oracle(fa, os.path.join(d, tag), args, seed=5)
        with api.GenReads(gamma=gamma, coverage=cov, layout=layout, seed=5) as g:
            g.load_profile(prof).load_genome(fa).create_frags().amplify()
            got = g.yield_reads_bytes()
            st = g.stats()
        exp = [H.read_bytes(p) for p in H.fastq_names(os.path.join(d, tag), layout)]
        assert list(got[:len(exp)]) == exp, tag
        return st, exp

    fa1 = os.path.join(d, "short.fa")
    write_fasta(fa1, [("chrA_1_900", synth_sequence(900, 1)), ("chrB_1_500", synth_sequence(500, 2))])
    st, exp = both(fa1, "short", "PE", 1e-8, 5.0)
    assert st["n_fulls"] == 0 and exp == [b"", b""]

    fa2 = os.path.join(d, "g.fa")
    write_fasta(fa2, [("chrA_1_60000", synth_sequence(60000, 3))])
    st, exp = both(fa2, "nocov", "SE", 2e-10, 0.001)
    assert st["reads_requested"] == 0 and exp == [b""]

    fa3 = os.path.join(d, "gap.fa")
    a, b = synth_sequence(40_000, 4), synth_sequence(30_011, 5)
    with open(fa3, "wb") as f:
        f.write(b">chrA_1_40000\n" + b"\n".join(a[i:i + 80].tobytes() for i in range(0, len(a), 80)) + b"\n>chrE_1_0\n>chrB_1_30011\n" +
                b"\n".join(b[i:i + 80].tobytes() for i in range(0, len(b), 80)))       # no newline at the end of the file
    st, exp = both(fa3, "gap", "PE", 5e-10, 6.0)
    assert st["n_fulls"] > 0 and len(exp[0]) > 10_000


# ---- the slab pipeline itself against the oracle: many slabs per file, so the double-buffered device slabs, the ring of pinned
# ---- host slots and the launch-ahead / finalize / consume hand-offs (reads.cu yield_reads) all cycle several times
@pytest.mark.parametrize("layout,profile,isize,slab", [("PE", "Illumina_HiSeq2500", 260, 1 << 20), ("SE", "Illumina_HiSeq2000", 260, 1 << 20),
                                                      ("PE", "Illumina_HiSeqXTen", 1200, 1 << 20), ("PE", "Illumina_HiSeq2500", 260, 64 << 10)])
def test_many_slabs_equal_the_oracle(tmp_path, layout, profile, isize, slab):
    cov = 40.0 if slab >= (1 << 20) else 4.0
    st = _run_case(str(tmp_path), f"slabs{layout}{isize}_{slab}", 1, 600_000, 11, profile, layout, 2e-10, cov, isize, seed=31337,
                   slab_bytes=slab, min_slabs=20)
    assert st["records"] > 0


def test_slabs_with_n_genome_equal_the_oracle(tmp_path):
    _run_case(str(tmp_path), "slabsN", 1, 300_000, 77, "Illumina_HiSeq2500", "PE", 3e-10, 30.0, 260, seed=9, with_n=True, slab_bytes=1 << 20, min_slabs=10)


# ---- the tables bench.py runs on (read length is a property of the .profile: the bench derives 150- / 100-bin profiles by
# ---- nearest-bin resampling; the oracle reads the same derived file). RL 150 is also the largest shared-memory footprint.
@pytest.mark.parametrize("src,rl,layout,isize", [("Illumina_HiSeq2500", 150, "PE", 260), ("Illumina_HiSeqXTen", 100, "PE", 260),
                                                 ("Illumina_HiSeq2500", 150, "SE", 260), ("Illumina_HiSeq2500", 250, "PE", 400)])   # 250: fewer warps per CTA
def test_resampled_bench_profiles_equal_the_oracle(tmp_path, src, rl, layout, isize):
    from scssim_b200.tools.resample_profile import resample
    prof = os.path.join(str(tmp_path), f"{src}_{rl}.profile")
    resample(H.profile_path(src), prof, rl)
    st = _run_case(str(tmp_path), f"res{rl}{layout}", 1, 400_000, 13, prof, layout, 2e-10, 6.0, isize, seed=150 + rl, slab_bytes=256 << 10, min_slabs=3)
    assert st["records"] > 0


def test_read_length_beyond_the_kernel_limit_is_a_clean_error(tmp_path):
    """Profiles whose quality tables do not fit shared memory (read length ~275+; the shipped profiles are 74-151) are refused."""
    from scssim_b200 import api
    from scssim_b200.synth import synth_sequence
    from scssim_b200.tools.resample_profile import resample
    prof = os.path.join(str(tmp_path), "p300.profile")
    resample(H.profile_path("Illumina_HiSeqXTen"), prof, 300)
    with api.GenReads(gamma=2e-10, coverage=4.0, isize=500, layout="PE", seed=3) as g:
        g.load_profile(prof).set_genome([("chrA_1_200000", synth_sequence(200_000, 5))]).create_frags().amplify()
        with pytest.raises(api.ScsError) as e:
            g.yield_reads_bytes()
        assert e.value.code == api.SCS_E_UNSUPPORTED


def test_slab_limits_are_clean_errors(tmp_path):
    """A slab that cannot hold 64 of the largest records is refused up front; a batch that outgrows its slab (a profile whose
    reads are much longer than the typical record the batch was sized for) ends with SCS_E_NOMEM — the compaction kernel skips
    what does not fit instead of writing past the slab — and the context stays usable with a larger slab."""
    from scssim_b200 import api
    from scssim_b200.synth import synth_sequence
    prof = H.profile_path("Illumina_HiSeq2500")
    seq = synth_sequence(300_000, 5)
    with api.GenReads(gamma=2e-10, coverage=8.0, layout="PE", seed=3, slab_bytes=1024) as g:
        g.load_profile(prof).set_genome([("chrA_1_300000", seq)]).create_frags().amplify()
        with pytest.raises(api.ScsError) as e:
            g.yield_reads_bytes()
        assert e.value.code == api.SCS_E_ARG
    lines = open(prof).read().split("\n")
    i = lines.index("[Insert Rate]")

    def with_insert_rate(rate, name):
        lines[i + 1] = str(rate)
        path = os.path.join(str(tmp_path), name)
        open(path, "w").write("\n".join(lines))
        return path
    # insertion rate 0.12: ~15 insertions per read, every record ~55 bytes longer than the typical record the batch is sized for
    with api.GenReads(gamma=2e-10, coverage=40.0, layout="PE", seed=3, slab_bytes=256 << 10) as g:
        g.load_profile(with_insert_rate(0.12, "heavier.profile")).set_genome([("chrA_1_300000", seq)]).create_frags().amplify()
        with pytest.raises(api.ScsError) as e:
            g.yield_reads_bytes()
        assert e.value.code == api.SCS_E_NOMEM, e.value
    heavy = with_insert_rate(0.06, "heavy.profile")   # ~7.5 insertions per read
    # same profile with a roomy slab: equals the oracle (reads with many indel events)
    _run_case(str(tmp_path), "heavy", 1, 300_000, 5, heavy, "PE", 2e-10, 8.0, 260, seed=3, slab_bytes=8 << 20)


@pytest.mark.parametrize("layout,direct", [("PE", True), ("PE", False), ("SE", True)])
def test_file_sink_equals_the_oracle(tmp_path, layout, direct, monkeypatch):
    """scs_yield_reads(prefix): the asynchronous file sink (ring of pinned slabs, writer threads, O_DIRECT with carried partial
    blocks or plain buffered writes) must leave exactly the oracle's files; 1 MiB slabs, so dozens of slabs per file."""
    from scssim_b200 import api
    if not direct:
        monkeypatch.setenv("SCS_NO_ODIRECT", "1")
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "cell.fa")
    H.write_genome(fa, 1, 500_000, 23)
    prof = H.profile_path("Illumina_HiSeq2500")
    args = H.genreads_args(prof, layout, 2e-10, 30.0, 260)
    H.run_oracle(fa, os.path.join(tmp, "orc"), args, seed=77)
    with api.GenReads(gamma=2e-10, coverage=30.0, layout=layout, seed=77, slab_bytes=1 << 20, ring_slabs=3, io_threads=3) as g:
        g.load_profile(prof).load_genome(fa).create_frags().amplify()
        want = g.plan_fastq_bytes()
        g.yield_reads(os.path.join(tmp, "gpu"))
        st = g.stats()
        g.yield_reads(os.path.join(tmp, "gpu"))          # a second run into the same files (truncate + rewrite)
    assert st["emit_launches"] >= 10
    for i, (a, b) in enumerate(zip(H.fastq_names(os.path.join(tmp, "gpu"), layout), H.fastq_names(os.path.join(tmp, "orc"), layout))):
        assert H.read_bytes(a) == H.read_bytes(b)
        assert want[i] == os.path.getsize(b)             # the sizing pass predicts the file size exactly


def _bgzf_members(data: bytes):
    """Walk a BGZF stream: yields (member bytes, uncompressed size); checks the BC subfield and BSIZE of every member."""
    import struct
    out, i = [], 0
    while i < len(data):
        assert data[i:i + 4] == b"\x1f\x8b\x08\x04" and data[i + 12:i + 16] == b"BC\x02\x00", i
        bsize = struct.unpack_from("<H", data, i + 16)[0] + 1
        assert bsize <= 65536 and i + bsize <= len(data)
        out.append((data[i:i + bsize], struct.unpack_from("<I", data, i + bsize - 4)[0]))
        i += bsize
    return out


@pytest.mark.parametrize("layout,with_n,slab", [("PE", False, 1 << 20), ("SE", True, 1 << 20), ("PE", False, 64 << 20)])
def test_gzip_output_decompresses_to_the_oracle(tmp_path, layout, with_n, slab):
    """gzip = 1: the FASTQ leaves the device as BGZF (independent gzip members of <= 32 KiB of text, dynamic-Huffman literals,
    CRC-32 combined across the CTA). Decompressed it must be the oracle's bytes; every member is well formed, none holds more
    than 32 KiB, the stream ends with the BGZF end-of-file marker; the same through the file sink (.fq.gz) and `gzip -dc`."""
    import gzip
    import subprocess
    from scssim_b200 import api
    tmp = str(tmp_path)
    fa = os.path.join(tmp, "cell.fa")
    H.write_genome_with_n(fa, 300_000, 77) if with_n else H.write_genome(fa, 1, 400_000, 29)
    prof = H.profile_path("Illumina_HiSeq2500")
    args = H.genreads_args(prof, layout, 3e-10, 20.0, 260)
    H.run_oracle(fa, os.path.join(tmp, "orc"), args, seed=5150)
    want = [H.read_bytes(p) for p in H.fastq_names(os.path.join(tmp, "orc"), layout)]
    with api.GenReads(gamma=3e-10, coverage=20.0, layout=layout, seed=5150, slab_bytes=slab, gzip=True, ring_slabs=3) as g:
        g.load_profile(prof).load_genome(fa).create_frags().amplify()
        got = g.yield_reads_bytes()
        st = g.stats()
        g.yield_reads(os.path.join(tmp, "gpu"))
    eof = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    for f, w in enumerate(want):
        assert gzip.decompress(got[f]) == w
        members = _bgzf_members(got[f])
        assert members[-1][0] == eof and max(n for _, n in members) <= 32768 and sum(n for _, n in members) == len(w)
        assert st["plain_bytes"][f] == len(w) and st["fastq_bytes"][f] == len(got[f]) and len(w) / len(got[f]) > 1.9
    names = [os.path.join(tmp, "gpu_1.fq.gz"), os.path.join(tmp, "gpu_2.fq.gz")] if layout == "PE" else [os.path.join(tmp, "gpu.fq.gz")]
    for p, w, gb in zip(names, want, got):
        assert H.read_bytes(p) == gb
        assert subprocess.run(["gzip", "-dc", p], capture_output=True, check=True).stdout == w


def test_gzip_of_an_empty_output_is_a_valid_stream(tmp_path):
    import gzip
    from scssim_b200 import api
    from scssim_b200.synth import synth_sequence
    prof = H.profile_path("Illumina_HiSeq2500")
    with api.GenReads(gamma=2e-10, coverage=0.001, layout="SE", seed=5, gzip=True) as g:
        g.load_profile(prof).set_genome([("chrA_1_60000", synth_sequence(60000, 3))]).create_frags().amplify()
        got = g.yield_reads_bytes()
    assert gzip.decompress(got[0]) == b""
