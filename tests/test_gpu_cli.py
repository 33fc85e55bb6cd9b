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
