"""Synthetic genome generator for tests and bench.py (SURVEY.md §8d "Synthetic inputs").

Per-block GC content ~ U[gc_lo, gc_hi] (100 kb blocks), bases i.i.d. within a block,
upper-case ACGT, fixed line width, sequence names ``<prefix><i>_<k>_<len>`` as the
reference's ``Malbac::yieldReads`` expects (/root/reference/lib/malbac/Malbac.cpp:413-419
parses the last ``_`` field as the haplotype length). Haplotype 2 = haplotype 1 + SNPs.
"""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def synth_sequence(length: int, seed: int, gc_lo: float = 0.35, gc_hi: float = 0.60,
                   block: int = 100_000) -> np.ndarray:
    """Return `length` ASCII bases (uint8) with a per-block GC spectrum."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty(length, dtype=np.uint8)
    nblk = (length + block - 1) // block
    gcs = rng.uniform(gc_lo, gc_hi, size=nblk)
    for b in range(nblk):
        lo, hi = b * block, min(length, (b + 1) * block)
        n = hi - lo
        is_gc = rng.random(n) < gcs[b]
        pick = rng.integers(0, 2, size=n, dtype=np.uint8)
        # A=0 C=1 G=2 T=3 ; GC -> {C,G}, AT -> {A,T}
        idx = np.where(is_gc, 1 + pick, 3 * pick)
        out[lo:hi] = _BASES[idx]
    return out


def add_snps(seq: np.ndarray, rate: float, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    out = seq.copy()
    n = int(len(seq) * rate)
    if n == 0:
        return out
    pos = rng.integers(0, len(seq), size=n)
    code = np.zeros(256, dtype=np.uint8)
    code[_BASES] = np.arange(4, dtype=np.uint8)
    shift = rng.integers(1, 4, size=n).astype(np.uint8)
    out[pos] = _BASES[(code[out[pos]] + shift) & 3]
    return out


def write_fasta(path: str, named_seqs, width: int = 100) -> None:
    """named_seqs: iterable of (name, uint8 array)."""
    with open(path, "wb") as f:
        for name, seq in named_seqs:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def synth_genome(n_chrom: int, chrom_len: int, seed: int, diploid: bool = True,
                 prefix: str = "chrS", snp_rate: float = 1e-3):
    """List of (name, seq) for a synthetic cell: per chromosome one or two haplotypes."""
    out = []
    for c in range(n_chrom):
        h1 = synth_sequence(chrom_len, seed * 1000 + c)
        out.append((f"{prefix}{c + 1}_1_{chrom_len}", h1))
        if diploid:
            h2 = add_snps(h1, snp_rate, seed * 1000 + 500 + c)
            out.append((f"{prefix}{c + 1}_2_{chrom_len}", h2))
    return out
