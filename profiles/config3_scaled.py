"""BASELINE configs[2] at 1/10 scale: MALBAC amplification of a synthetic haploid genome at the DEFAULT primer rate
(gamma 1e-9), then SE reads with the HiSeq2000 (75 bp) profile at 5x. Reports the amplicon tree and stage rates.
usage: python profiles/config3_scaled.py [genome_len]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from scssim_b200 import api
from scssim_b200.synth import synth_sequence

glen = int(sys.argv[1]) if len(sys.argv) > 1 else 310_000_000
seq = synth_sequence(glen, 4242)
with api.GenReads(gamma=1e-9, coverage=10.0, layout="SE", seed=3, slab_bytes=64 << 20) as g:
    g.load_profile(H.profile_path("Illumina_HiSeq2000"))
    t0 = time.time(); g.set_genome([(f"chrS1_1_{glen}", seq)]).create_frags(); t1 = time.time()
    g.amplify(); t2 = time.time()
    g.set_read_counts(); t3 = time.time()
    g.yield_reads_discard(); t4 = time.time()
    st = g.stats()
out = {"genome_len": glen, "gamma": 1e-9, "frags": st["n_frags"], "semis": st["n_semis"], "fulls": st["n_fulls"], "records": st["records"],
       "ms": {"pack+frags": (t1 - t0) * 1e3, "amplify": st["ms_amplify"], "alloc": st["ms_alloc"], "reads": st["ms_reads"]},
       "amplicons_per_s": (st["n_semis"] + st["n_fulls"]) / (st["ms_amplify"] / 1e3),
       "reads_per_s": st["records"] / (st["ms_reads"] / 1e3), "fastq_bytes": st["fastq_bytes"], "primers_left": st["total_primers_left"]}
print(json.dumps(out))
