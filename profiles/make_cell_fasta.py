"""Write the bench's synthetic cell (bench.py: 4 chromosomes x 2 haplotypes, GC 35-60 %, haplotype 2 = haplotype 1 + 0.1 % SNPs) as a FASTA
file for CLI runs: python profiles/make_cell_fasta.py <scale> <out.fa>   (scale 1.0 = 6.2 Gb)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from scssim_b200.synth import write_fasta

scale, out = float(sys.argv[1]), sys.argv[2]
clen = int(bench.CHROM_LEN * scale)
seqs = []
for c in range(bench.N_CHROM):
    b1 = torch.empty(clen, dtype=torch.uint8, pin_memory=True).numpy(); b2 = torch.empty(clen, dtype=torch.uint8, pin_memory=True).numpy()
    bench.synth_chromosome_cuda(torch, clen, 7000 + c, "cuda:0", b1, b2)
    seqs += [(f"chrS{c + 1}_1_{clen}", b1), (f"chrS{c + 1}_2_{clen}", b2)]
write_fasta(out, seqs)
bench_profile_dir = os.path.dirname(out)
print(bench.bench_profile(bench_profile_dir))
