#!/usr/bin/env python
"""bench.py — the `scssim genreads` hot path on B200 (contract: see the task brief; one JSON line on stdout).

Workload at N=1 = BASELINE.json configs[1]: synthetic 250 Mb haploid sequence (GC 35-60 % spectrum), paired-end
150 bp at 10x with the HiSeq2500 profile (resampled 125 -> 150 bins, scssim_b200/tools/resample_profile.py; read
length is a property of the .profile in the reference), GC bias on, gamma 2e-10 (README value). The reference's
`-c` is relative to HALF the summed sequence length and counts individual reads (Malbac.cpp:414-420), so 10x of a
haploid FASTA is `-c 20`: 16.67 M reads = 8.33 M pairs per step.

A "step" = one pass of the hot path over that input: MALBAC amplification -> GC-weighted read allocation -> read
synthesis -> FASTQ packing, FASTQ landing in the library's pinned host ring.
  value : M reads/s with the packed genome already resident in HBM when the timed region starts (CUDA events).
  e2e   : the same through the C-ABI calls a user makes with HOST buffers: scs_set_genome (H2D of the ASCII genome
          from pinned memory + pack) ... scs_yield_reads_sink (D2H of every FASTQ byte) inside the timed region.
For N > 1 (torchrun) the cell has N such chromosomes, one per rank (weak scaling: per-GPU work fixed). The ranks form ONE
run: NCCL all-reduces carry the primer budget per amplification round, the per-batch amplicon counts and the cell-wide
weight vector of the read allocation; every rank then writes its own FASTQ shard.

--impl reference times the reference's own CPU implementation (oracle/_ref/bin/scssim, built from /root/reference by
oracle/build_ref.sh) with all host threads on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GENOME_LEN = 250_000_000
COVERAGE = 20.0          # = 10x of a haploid FASTA (see module docstring)
GAMMA = 2e-10
READ_LEN = 150
SAMPLE_DIV = 2           # CPU legs run on a 1/2-scale genome (125 Mb, 8.3 M reads), same gamma / coverage / profile: ~20 s of reference work
UNIT = "M reads/s"
METRIC = "PE150 M reads/s (FASTQ GB/s in config) vs HBM/D2H roofline"


def bench_profile(tmp):
    import helpers as H
    from scssim_b200.tools.resample_profile import resample
    p = os.path.join(tmp, "HiSeq2500_150.profile")
    resample(H.profile_path("Illumina_HiSeq2500"), p, READ_LEN)
    return p


def workload_name():
    return ("synthetic 250 Mb haploid (GC 35-60%), PE150 10x (-c 20, 16.67 M reads/step), HiSeq2500 profile resampled to "
            "150 bins, GC bias on, gamma 2e-10")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference_cpu(tmp, profile, steps, warmup, genome_len=GENOME_LEN):
    """Reference genreads on the host cores, bounded sample. Returns (M reads/s, cores, sample text, s/step)."""
    import helpers as H
    from scssim_b200.synth import synth_sequence, write_fasta
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "scssim")
    cores = os.cpu_count() or 1
    n = genome_len // SAMPLE_DIV
    fa = os.path.join(tmp, "sample.fa")
    write_fasta(fa, [(f"chrS1_1_{n}", synth_sequence(n, 7001))])
    reads = int((n // 2) * COVERAGE / READ_LEN)
    sample = (f"1/{SAMPLE_DIV}-scale genome ({n / 1e6:.1f} Mb haploid, {reads / 1e6:.2f} M reads/step), same gamma/coverage/profile; "
              f"reference `scssim genreads -t {cores}` whole-process wall time (load+amplify+reads+FASTQ files on local disk)")
    if not os.path.exists(exe):
        # no compiled reference on this box: time the CPU oracle (single thread) instead
        kind, cores = "port", 1
        cmd = [H.oracle_bin(), "genreads", "-i", fa, "-o", os.path.join(tmp, "cpu"), "--seed", "1"] + H.genreads_args(profile, "PE", GAMMA, COVERAGE, 260)
        sample = sample.replace(f"reference `scssim genreads -t {os.cpu_count() or 1}`", "CPU oracle (oracle/scs_oracle, 1 thread)")
    else:
        kind = "reference"
        cmd = [exe, "genreads", "-i", fa, "-t", str(cores), "-o", os.path.join(tmp, "cpu")] + H.genreads_args(profile, "PE", GAMMA, COVERAGE, 260)
    times = []
    for i in range(warmup + steps):
        if os.path.exists(fa + ".fai"):
            os.remove(fa + ".fai")
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return reads / sec / 1e6, cores, sample, sec, kind, reads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-len", type=int, default=GENOME_LEN, help="(debug) override the workload size; invalidates the number")
    ap.add_argument("--coverage", type=float, default=COVERAGE, help="(debug) override -c; invalidates the number")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="(debug) skip the CPU leg")
    ap.add_argument("--no-balance", action="store_true", help="(debug) N > 1: every rank writes the reads of its own amplicons (equal shards)")
    ap.add_argument("--slab-mb", type=int, default=64, help="FASTQ staging slab per file and buffer (MiB)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(a.warmup, 0)
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if a.impl == "reference":
        if rank != 0:
            return 0
        with tempfile.TemporaryDirectory() as tmp:
            profile = bench_profile(tmp)
            v, cores, sample, sec, kind, reads = run_reference_cpu(tmp, profile, max(1, min(a.steps, 3)), min(warmup, 1), a.genome_len)
        emit({"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": workload_name() if a.genome_len == GENOME_LEN else f"DEBUG {a.genome_len} bp", "sample": sample},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from scssim_b200 import api
    from scssim_b200.synth import synth_sequence

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with tempfile.TemporaryDirectory() as tmp:
        profile = bench_profile(tmp)
        glen = a.genome_len
        # every rank: its own chromosome of the synthetic cell (weak scaling), ASCII in pinned host memory
        seq = synth_sequence(glen, 7000 + rank)
        pinned = torch.empty(glen, dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = seq
        named = [(f"chrS{rank + 1}_1_{glen}", pinned.numpy())]
        # pinned D2H copy ceiling of this box, measured live with ALL ranks copying at once (the binding roof of the read
        # stage; on shared PCIe fabrics the per-GPU rate drops as N grows)
        nb = 256 << 20
        dbuf = torch.empty(nb, dtype=torch.uint8, device="cuda"); hbuf = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
        hbuf.copy_(dbuf, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(6):
            hbuf.copy_(dbuf, non_blocking=True)
        c1.record(); torch.cuda.synchronize()
        mine = torch.tensor([6 * nb / (c0.elapsed_time(c1) / 1e3) / 1e9], dtype=torch.float64, device="cuda")
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(per_rank, mine)
        else:
            per_rank = [mine]
        d2h_per_rank = [float(x) for x in per_rank]
        d2h_peak = sum(d2h_per_rank)            # aggregate over the N GPUs
        d2h_slowest = min(d2h_per_rank)
        del dbuf, hbuf
        # one cell sharded over the ranks: same seed everywhere, global ids key every Philox stream. With several GPUs the
        # read slots are cut in proportion to each GPU's measured D2H rate (balance=1), so a GPU behind a slower host link
        # writes fewer reads; the rank-ordered shards are byte-identical to the single-GPU files
        g = api.GenReads(gamma=GAMMA, coverage=a.coverage, isize=260, layout="PE", seed=0x5C55, device=local, rank=rank, world=world,
                         slab_bytes=a.slab_mb << 20, balance=(world > 1 and not a.no_balance))
        if world > 1:
            g.set_shard_weight(d2h_per_rank[rank])
        if world > 1:
            from scssim_b200.dist import make_collectives, make_device_allreduce
            g.set_collectives(*make_collectives(dist, device=f"cuda:{local}"))
            g.set_device_collective(*make_device_allreduce(dist, f"cuda:{local}"))
        g.load_profile(profile)
        g.set_genome(named).create_frags()

        def step_resident():
            g.amplify().set_read_counts().yield_reads_discard()

        sink_state = {"bytes": 0, "sum": 0}

        def sink(_u, f, data, n):
            sink_state["bytes"] += n
            sink_state["sum"] += C.cast(data, C.POINTER(C.c_ubyte))[0]   # touch the landed bytes
            return 0
        sink_cb = api.SINK_FN(sink)

        def step_e2e():
            g.set_genome(named).create_frags().amplify().set_read_counts()
            g._ck(api.lib().scs_yield_reads_sink(g._h, sink_cb, None))

        for _ in range(warmup):
            step_resident()
        l0 = g.stats()["kernel_launches"]
        clocks = ClockSampler(local); clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        emit_ms = emit_n = reads_ms = amp_ms = alloc_ms = 0.0
        barrier(); ev0.record(); t0 = time.perf_counter()
        for _ in range(a.steps):
            step_resident()
            st = g.stats()
            emit_ms += st["ms_emit_kernel"]; emit_n += st["emit_launches"]; reads_ms += st["ms_reads"]; amp_ms += st["ms_amplify"]; alloc_ms += st["ms_alloc"]
        ev1.record(); barrier(); wall = time.perf_counter() - t0
        # the library works on its own streams: the torch events bracket host-synchronous calls, so use the larger of the two clocks
        dev_s = max(ev0.elapsed_time(ev1) / 1e3, wall)
        clk = clocks.stop()
        st = g.stats()
        launches = st["kernel_launches"] - l0
        reads_per_step = st["records"]   # FASTQ records actually written (both files)
        fastq_bytes = st["fastq_bytes"][0] + st["fastq_bytes"][1]
        n_fulls, n_semis = st["n_fulls"], st["n_semis"]

        # e2e leg (host buffers in, host bytes out); two untimed passes first (pool growth, page faults of the host side)
        step_e2e()
        step_e2e()
        barrier(); t0 = time.perf_counter()
        e2e_steps = max(1, min(a.steps, 3))
        for _ in range(e2e_steps):
            step_e2e()
        barrier(); e2e_s = time.perf_counter() - t0

        tmax = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
        tot = torch.tensor([float(reads_per_step), float(fastq_bytes)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dev_s, e2e_s = tmax.tolist(); reads_all, bytes_all = tot.tolist()
        value = reads_all * a.steps / dev_s / 1e6
        e2e_value = reads_all * e2e_steps / e2e_s / 1e6

        cpu = None
        if rank == 0 and not a.no_cpu_baseline:
            v, cores, sample, sec, kind, _ = run_reference_cpu(tmp, profile, 1, 0, glen)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "s_per_sample": sec}
        g.close()

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        # algorithmic HBM bytes of one emit launch (DESIGN.md "emit_kernel"): FASTQ bytes stored (into the staging buffer) + 2-bit
        # genome windows read (ceil(L/4) B per read) + 16 B descriptor per amplicon touched + 8 B of record sizes per slot
        slots = reads_per_step / 2
        alg_step = fastq_bytes + reads_per_step * ((READ_LEN + 3) // 4) + 16 * n_fulls + 8 * slots
        traffic = None
        try:   # dram__bytes_read+write of one emit launch from the committed ncu --set full capture of this same command
            with open(os.path.join(ROOT, "profiles", "emit_traffic.json")) as f:
                traffic = json.load(f)["dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        emit_launches_per_step = emit_n / a.steps if a.steps else 0
        emit_ms_avg = emit_ms / emit_n if emit_n else None
        achieved = (alg_step / emit_launches_per_step) / (emit_ms_avg / 1e3) / 1e9 if emit_ms_avg else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_s / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name() if (glen == GENOME_LEN and a.coverage == COVERAGE) else f"DEBUG {glen} bp per rank, -c {a.coverage}", "per_gpu": "one 250 Mb chromosome per rank; ranks form one cell (global read allocation over NCCL)", "layout": "PE", "read_length": READ_LEN,
                       "reads_per_step": reads_all, "fastq_bytes_per_step": bytes_all, "fastq_GBps": bytes_all * a.steps / dev_s / 1e9,
                       "full_amplicons": n_fulls, "semi_amplicons": n_semis,
                       "l2": "every step streams ~5 GB of FASTQ through L2 (>> 126 MB), evicting the 62 MB packed genome between steps",
                       "stage_ms_per_step": {"amplify": amp_ms / a.steps, "alloc": alloc_ms / a.steps, "reads": reads_ms / a.steps}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": glen * world, "d2h_bytes_per_step": bytes_all, "steps": e2e_steps,
                    "fastq_GBps": bytes_all * e2e_steps / e2e_s / 1e9},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "emit_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "algorithmic_bytes_per_launch": (alg_step / emit_launches_per_step) if emit_launches_per_step else None, "peak_source": peak_src, "launches_per_step": emit_launches_per_step, "avg_launch_ms": emit_ms_avg,
                         "kernel_share_of_step": (emit_ms / a.steps) / (dev_s / a.steps * 1e3) if dev_s else None,
                         "d2h": {"bound": "pcie", "achieved": bytes_all * a.steps / dev_s / 1e9, "peak": d2h_peak, "unit": "GB/s (all GPUs)",
                                 "frac": (bytes_all * a.steps / dev_s / 1e9 / d2h_peak) if d2h_peak else None,
                                 "frac_read_stage_rank0": (bytes_all / world / (reads_ms / a.steps / 1e3) / 1e9 / d2h_per_rank[0]) if reads_ms else None,
                                 "per_gpu_peak": [round(x, 1) for x in d2h_per_rank],
                                 "equal_shard_ceiling": world * d2h_slowest, "balanced": bool(world > 1 and not a.no_balance),
                                 "note": "FASTQ bytes landing in pinned host memory over the whole step, against the pinned D2H copy rate measured in this run with all ranks copying at once; this, not HBM, is the binding roof of the path. With equal shards the slowest GPU's link would set the pace (N x slowest); for N > 1 the read slots are cut in proportion to the measured per-GPU rates"}},
            "cpu_baseline": cpu, "clocks": clk,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
