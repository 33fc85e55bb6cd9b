"""One GPU's share of BASELINE configs[3] (6.2 Gb diploid cell, PE150 30x = `-c 60`, 8 GPUs): 775 Mb of sequence
(3 synthetic chromosomes), 155 M reads (77.5 M pairs), ~49 GB of FASTQ landing in the pinned host ring.
North-star target: the whole cell (8 such shares in parallel) in < 30 s. usage: python profiles/config4_share.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import bench
import helpers as H
from scssim_b200 import api
from scssim_b200.synth import synth_sequence
import tempfile

n_seq, slen = 3, 258_333_333
t0 = time.time()
named = [(f"chrS{i + 1}_1_{slen}", synth_sequence(slen, 9000 + i)) for i in range(n_seq)]
t_synth = time.time() - t0
with tempfile.TemporaryDirectory() as tmp:
    prof = bench.bench_profile(tmp)
    with api.GenReads(gamma=2e-10, coverage=60.0, layout="PE", seed=11, slab_bytes=64 << 20) as g:
        g.load_profile(prof)
        t0 = time.time(); g.set_genome(named).create_frags(); t1 = time.time()
        g.amplify().set_read_counts(); t2 = time.time()
        g.yield_reads_discard(); t3 = time.time()
        st = g.stats()
dev_s = (st["ms_amplify"] + st["ms_alloc"] + st["ms_reads"]) / 1e3
print(json.dumps({"share_bases": n_seq * slen, "reads": st["records"], "fastq_GB": sum(st["fastq_bytes"]) / 1e9, "full_amplicons": st["n_fulls"],
                  "ms": {"set_genome+frags": (t1 - t0) * 1e3, "amplify": st["ms_amplify"], "alloc": st["ms_alloc"], "reads": st["ms_reads"]},
                  "device_seconds": dev_s, "M_reads_per_s": st["records"] / dev_s / 1e6, "fastq_GBps": sum(st["fastq_bytes"]) / dev_s / 1e9,
                  "host_synth_s": t_synth}))
