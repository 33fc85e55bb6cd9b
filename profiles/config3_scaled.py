"""BASELINE configs[2]: MALBAC amplification of a synthetic haploid genome at the DEFAULT primer rate (gamma 1e-9), then SE
reads with the HiSeq2000 (75 bp) profile at 5x (-c 10 of a haploid FASTA). Reports the amplicon tree and stage rates.
usage: python profiles/config3_scaled.py [sequence_len] [n_sequences]     (default: 1/10 scale, one 310 Mb sequence;
       full scale = 1550000000 2, two sequences because one sequence must stay below 2^31 bases, lib/fastahack/Fasta.h:36)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from scssim_b200 import api
from scssim_b200.synth import synth_sequence

glen = int(sys.argv[1]) if len(sys.argv) > 1 else 310_000_000
nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t_s = time.time()
named = [(f"chrS{i + 1}_1_{glen}", synth_sequence(glen, 4242 + i)) for i in range(nseq)]
t_synth = time.time() - t_s
with api.GenReads(gamma=1e-9, coverage=10.0, layout="SE", seed=3, slab_bytes=int(os.environ.get("SLAB_MB", "64")) << 20) as g:
    g.load_profile(H.profile_path("Illumina_HiSeq2000"))
    t0 = time.time(); g.set_genome(named).create_frags(); t1 = time.time()
    g.amplify(); t2 = time.time()
    g.set_read_counts(); t3 = time.time()
    g.yield_reads_discard(); t4 = time.time()
    st = g.stats()
out = {"genome_len": glen * nseq, "n_sequences": nseq, "host_synth_s": t_synth, "wall_s": {"amplify": t2 - t1, "alloc": t3 - t2, "reads": t4 - t3},
        "gamma": 1e-9, "frags": st["n_frags"], "semis": st["n_semis"], "fulls": st["n_fulls"], "records": st["records"],
       "ms": {"pack+frags": (t1 - t0) * 1e3, "amplify": st["ms_amplify"], "alloc": st["ms_alloc"], "reads": st["ms_reads"]},
       "amplicons_per_s": (st["n_semis"] + st["n_fulls"]) / (st["ms_amplify"] / 1e3),
       "reads_per_s": st["records"] / (st["ms_reads"] / 1e3), "ms_reads_kernels": st["ms_reads_kernels"], "ms_emit_kernel": st["ms_emit_kernel"], "emit_launches": st["emit_launches"], "fastq_bytes": st["fastq_bytes"], "primers_left": st["total_primers_left"]}
print(json.dumps(out))
