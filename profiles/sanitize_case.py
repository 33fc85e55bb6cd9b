"""Smallest paired-end case for compute-sanitizer (one tool per gpurun call): 150 kb diploid chromosome, default gamma for the
amplification kernels, PE at 8x through 64 KiB slabs (≈ 40 slabs per file: the double-buffered device slabs, the pinned ring
and both streams all cycle), then the same through the asynchronous file sink, and a small simuvars run.
usage: compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python profiles/sanitize_case.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from scssim_b200 import api
from scssim_b200.synth import synth_genome

prof = H.profile_path("Illumina_HiSeq2500")
genome = synth_genome(1, 150_000, seed=5, diploid=True)
with tempfile.TemporaryDirectory() as tmp:
    with api.GenReads(gamma=1e-9, coverage=8.0, layout="PE", seed=7, slab_bytes=64 << 10, ring_slabs=3, io_threads=2) as g:
        g.load_profile(prof).set_genome(genome).create_frags().amplify().set_read_counts()
        f1, f2 = g.yield_reads_bytes()
        st = g.stats()
        g.yield_reads(os.path.join(tmp, "out"))
        same = open(os.path.join(tmp, "out_1.fq"), "rb").read() == f1 and open(os.path.join(tmp, "out_2.fq"), "rb").read() == f2
    print(f"sanitize_case: {st['n_fulls']} full amplicons, {st['records']} records, {st['emit_launches']} slabs, files == sink bytes: {same}")
    ref, snp, var = H.make_simuvars_case(os.path.join(tmp, "sv"), 3, chrom_lens=(40_000, 20_000))
    with api.GenReads() as g:
        n = len(g.simuvars_bytes(ref, snp, var))
    print(f"sanitize_case: simuvars {n} FASTA bytes")
    sys.exit(0 if same and st["emit_launches"] >= 10 else 1)
