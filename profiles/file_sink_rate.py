"""FASTQ to FILES: `scs_yield_reads(prefix)` with 1..16 writer threads on this box, page cache and tmpfs.
One JSON line. Usage: python profiles/file_sink_rate.py [genome_Mb]"""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H                                        # noqa: E402
from scssim_b200 import api                                # noqa: E402
from scssim_b200.synth import synth_sequence               # noqa: E402
from scssim_b200.tools.resample_profile import resample    # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
out = {"genome_mb": mb, "runs": []}
with tempfile.TemporaryDirectory() as tmp:
    prof = os.path.join(tmp, "p150.profile")
    resample(H.profile_path("Illumina_HiSeq2500"), prof, 150)
    seq = synth_sequence(mb * 1_000_000, 7000)
    named = [(f"chrS1_1_{len(seq)}", seq)]
    for where in ("/tmp", "/dev/shm"):
        for threads in (1, 2, 4, 8, 16):
            d = tempfile.mkdtemp(dir=where)
            try:
                with api.GenReads(gamma=2e-10, coverage=20.0, layout="PE", seed=0x5C55, io_threads=threads) as g:
                    g.load_profile(prof).set_genome(named).create_frags().amplify().set_read_counts()
                    g.yield_reads_discard()                                   # warm the slabs
                    t0 = time.perf_counter()
                    g.yield_reads(os.path.join(d, "reads"))
                    dt = time.perf_counter() - t0
                    st = g.stats()
                nbytes = sum(st["fastq_bytes"])
                assert os.path.getsize(os.path.join(d, "reads_1.fq")) == st["fastq_bytes"][0]
                out["runs"].append({"where": where, "threads": threads, "s": dt, "GBps": nbytes / dt / 1e9, "M_reads_per_s": st["records"] / dt / 1e6,
                                    "ms_reads_device": st["ms_reads"]})
            finally:
                shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out))
