"""Shared test plumbing: fixtures on disk, the CPU oracle (oracle/), the compiled reference (oracle/_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/ — the product
(scssim_b200/) never does.
"""
import ctypes as C
import functools
import lzma
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
PROFILES = ["Illumina_HiSeq2500", "Illumina_HiSeqXTen", "Illumina_HiSeq2000", "Illumina_GenomeAnalyzerIIx"]


@functools.lru_cache(maxsize=None)
def scratch_dir() -> str:
    d = os.environ.get("SCS_TEST_SCRATCH", "/tmp/scssim_b200_tests")
    os.makedirs(d, exist_ok=True)
    return d


@functools.lru_cache(maxsize=None)
def profile_path(name: str) -> str:
    """Decompress the committed copy of a shipped .profile (test fixture) and return its path."""
    out = os.path.join(scratch_dir(), name + ".profile")
    if not os.path.exists(out):
        tmp = f"{out}.{os.getpid()}.tmp"      # several ranks may decompress at once: unique name, atomic rename
        with lzma.open(os.path.join(GOLDEN, "profiles", name + ".profile.xz")) as f, open(tmp, "wb") as o:
            o.write(f.read())
        os.replace(tmp, out)
    return out


@functools.lru_cache(maxsize=None)
def oracle_bin() -> str:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "scs_oracle", "liboracle.so"], check=True)
    return os.path.join(ORACLE_DIR, "scs_oracle")


@functools.lru_cache(maxsize=None)
def oracle_lib():
    oracle_bin()
    L = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
    L.orc_profile_load.restype = C.c_void_p
    L.orc_profile_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.orc_profile_free.argtypes = [C.c_void_p]
    L.orc_profile_info.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_profile_scalars.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_profile_cdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_predict.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                              C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_det_log.restype = C.c_double
    L.orc_det_log.argtypes = [C.c_double]
    L.orc_philox_draw.restype = C.c_uint32
    L.orc_philox_draw.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
    return L


def ref_replay_bin():
    """The compiled reference with seed/draw-log hooks (oracle/build_ref.sh); None if it was never built."""
    p = os.path.join(ORACLE_DIR, "_ref", "bin", "scssim_replay")
    if os.path.exists(p):
        return p
    if os.path.isdir("/root/reference/lib"):
        subprocess.run(["bash", os.path.join(ORACLE_DIR, "build_ref.sh")], check=True)
        return p if os.path.exists(p) else None
    return None


def write_genome(path: str, n_chrom: int, chrom_len: int, seed: int, diploid: bool = True, width: int = 100):
    from scssim_b200.synth import synth_genome, write_fasta
    g = synth_genome(n_chrom, chrom_len, seed, diploid=diploid)
    write_fasta(path, g, width=width)
    return g


def genreads_args(profile, layout="PE", gamma=2e-10, coverage=5.0, isize=260, primers=100000):
    return ["-m", profile, "-l", layout, "-r", repr(gamma), "-c", repr(coverage), "-s", str(isize), "-p", str(primers)]


def run_reference_replay(fa: str, out_prefix: str, tape_prefix: str, seed: int, args):
    """Run the reference itself (-t 1) with a fixed seed, logging every draw. Returns FASTQ paths."""
    exe = ref_replay_bin()
    assert exe is not None
    if os.path.exists(fa + ".fai"):
        os.remove(fa + ".fai")
    env = dict(os.environ, SCS_SEED=str(seed), SCS_REPLAY_LOG=tape_prefix)
    subprocess.run([exe, "genreads", "-i", fa, "-t", "1", "-o", out_prefix] + list(args), check=True, env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def run_oracle(fa: str, out_prefix: str, args, tape_prefix=None, seed=None, dump_prefix=None):
    cmd = [oracle_bin(), "genreads", "-i", fa, "-o", out_prefix] + list(args)
    if tape_prefix is not None:
        cmd += ["--tape", tape_prefix]
    else:
        cmd += ["--seed", str(seed)]
    if dump_prefix is not None:
        cmd += ["--dump", dump_prefix]
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    return r.stderr.decode()


def read_bytes(path):
    with open(path, "rb") as f:
        return f.read()


def fastq_names(prefix, layout):
    return [prefix + "_1.fq", prefix + "_2.fq"] if layout == "PE" else [prefix + ".fq"]


def oracle_dump(dump_prefix):
    d = {}
    d["frags"] = np.fromfile(dump_prefix + ".frags.i64", dtype=np.int64).reshape(-1, 5)
    d["semis"] = np.fromfile(dump_prefix + ".semis.u32", dtype=np.uint32).reshape(-1, 7)
    d["fulls"] = np.fromfile(dump_prefix + ".fulls.u32", dtype=np.uint32).reshape(-1, 7)
    d["errs"] = np.fromfile(dump_prefix + ".errs.u32", dtype=np.uint32).reshape(-1, 2)
    d["counts"] = np.fromfile(dump_prefix + ".counts.u32", dtype=np.uint32)
    d["weights"] = np.fromfile(dump_prefix + ".weights.f64", dtype=np.float64)
    d["primer_counts"] = np.fromfile(dump_prefix + ".primer_counts.i64", dtype=np.int64)
    d["meta"] = np.fromfile(dump_prefix + ".meta.u64", dtype=np.uint64)
    return d


def expected_windows(genome, od):
    """Oriented-window form (gstart, rc, len) of the oracle's semi and full amplicons, from its
    (fragment, spos, len) records — the representation the CUDA path stores."""
    goff, acc = [], 0
    for _, s in genome:
        goff.append(acc)
        acc += (len(s) + 31) // 32 * 32
    fr = od["frags"]
    g = np.array([goff[int(s)] for s in fr[:, 0]], dtype=np.int64) + fr[:, 1]
    L = fr[:, 2]
    t_rc = (fr[:, 3] == 1)
    t_g = np.where(t_rc, g + L - 1, g)
    se = od["semis"].astype(np.int64)
    f = se[:, 0]
    s, l = se[:, 1], se[:, 2]
    u_rc = ~t_rc[f]
    u_g = np.where(t_rc[f], t_g[f] - s - l + 1, t_g[f] + s + l - 1)
    fu = od["fulls"].astype(np.int64)
    p = fu[:, 0]
    s2, l2 = fu[:, 1], fu[:, 2]
    f_rc = u_rc[p]
    f_g = np.where(f_rc, u_g[p] - s2, u_g[p] + s2)
    semis = np.stack([u_g, u_rc.astype(np.int64), l, se[:, 3], se[:, 4]], axis=1)
    fulls = np.stack([f_g, f_rc.astype(np.int64), l2, fu[:, 3]], axis=1)
    return semis, fulls


def write_genome_with_n(path: str, glen: int = 300_000, seed: int = 77, width: int = 70):
    """Synthetic diploid chromosome with an N run, scattered N, IUPAC codes and soft-masked (lower-case) stretches."""
    from scssim_b200.synth import synth_genome, write_fasta
    g = synth_genome(1, glen, seed, diploid=True)
    rng = np.random.default_rng(seed)
    out = []
    for name, s in g:
        s = s.copy()
        s[glen * 2 // 5:glen * 2 // 5 + 5000] = ord("N")
        s[rng.integers(0, glen, 120)] = ord("N")
        s[rng.integers(0, glen, 10)] = ord("R")
        for lo in rng.integers(0, glen - 2000, 5):
            s[lo:lo + 1500] |= 0x20          # lower case
        out.append((name, s))
    write_fasta(path, out, width=width)
    return out
