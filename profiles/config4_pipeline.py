"""BASELINE configs[4], one GPU's share (or the whole thing under torchrun): tumour-like cell — `simuvars` with the variant
density of 3 M SNPs / 500 k indels / 2 k CNVs per 3.1 Gb on a synthetic reference, the simulated diploid cell going STRAIGHT
into the genreads genome on the device (scs_simuvars_to_genome, no FASTA round trip), then genreads PE100 at 60x with the
HiSeqXTen profile resampled to 100 bins.
Coverage convention (Malbac.cpp:414-420): reads = (sum of the `_<len>` name suffixes / 2) * c / RL individual reads; the
haplotype names carry the REFERENCE chromosome length, so 60x of the diploid cell = `-c 60`.
usage: python profiles/config4_pipeline.py [reference_Mb_per_rank] [n_chrom_per_rank]      (torchrun for N > 1)"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "profiles"))
import helpers as H                                        # noqa: E402
from scssim_b200 import api                                # noqa: E402
from scssim_b200.tools.resample_profile import resample    # noqa: E402
from simuvars_scale import make_inputs                     # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 388
nchr = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tmp = os.environ.get("SCS_C4_DIR") or os.path.join(tempfile.gettempdir(), "scs_config4")
os.makedirs(tmp, exist_ok=True)
t0 = time.time()
if rank == 0:
    ref, snp, var, meta = make_inputs(tmp, mb * world, nchr * world)
    json.dump(meta, open(os.path.join(tmp, "meta.json"), "w"))
if dist:
    dist.barrier()
ref, snp, var = (os.path.join(tmp, f) for f in ("ref.fa", "snp.txt", "vars.txt"))
meta = json.load(open(os.path.join(tmp, "meta.json")))
t_inputs = time.time() - t0
prof = os.path.join(tmp, f"xten100.{rank}.profile")
resample(H.profile_path("Illumina_HiSeqXTen"), prof, 100)
g = api.GenReads(gamma=2e-10, coverage=60.0, layout="PE", seed=0xC4, device=local, rank=rank, world=world)
if dist:
    from scssim_b200.dist import make_collectives, make_device_allreduce
    g.set_collectives(*make_collectives(dist, device=f"cuda:{local}"))
    g.set_device_collective(*make_device_allreduce(dist, f"cuda:{local}"))
g.load_profile(prof)
out = {"world": world, "reference_mb": mb * world, "inputs": meta, "host_input_synthesis_s": t_inputs}
for it in range(2):          # second pass = warm (page cache, CUDA pools)
    if dist:
        dist.barrier()
    t0 = time.perf_counter(); g.simuvars_to_genome(ref, snp, var); t1 = time.perf_counter()
    sv = g.simuvars_stats()
    g.create_frags(); t2 = time.perf_counter()
    g.amplify().set_read_counts(); t3 = time.perf_counter()
    g.yield_reads_discard(); t4 = time.perf_counter()
    st = g.stats()
    out["pass%d" % it] = {"simuvars_to_genome_s": t1 - t0, "simuvars": {k: sv[k] for k in ("ms_read", "ms_plan", "ms_device", "ms_kernels", "n_pieces", "n_subs", "n_segments", "out_bases")},
                          "create_frags_s": t2 - t1, "amplify_alloc_s": t3 - t2, "reads_s": t4 - t3, "total_s": t4 - t0,
                          "cell_bases_this_rank": st["genome_bases"], "records_this_rank": st["records"], "fastq_GB_this_rank": sum(st["fastq_bytes"]) / 1e9,
                          "full_amplicons_this_rank": st["n_fulls"], "device_ms": {"amplify": st["ms_amplify"], "alloc": st["ms_alloc"], "reads": st["ms_reads"]},
                          "M_reads_per_s_this_rank": st["records"] / (t4 - t1) / 1e6}
g.close()
if dist:
    import torch
    v = torch.tensor([out["pass1"]["total_s"], -float(out["pass1"]["records_this_rank"])], dtype=torch.float64, device="cuda")
    mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = v.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    out["all_ranks"] = {"total_s_max": float(mx[0]), "records": -float(sm[1]), "M_reads_per_s": -float(sm[1]) / float(mx[0]) / 1e6}
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps(out))
