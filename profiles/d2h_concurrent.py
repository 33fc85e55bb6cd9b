"""Aggregate pinned D2H bandwidth with all ranks copying at once (torchrun): the platform ceiling for FASTQ landing in host
memory at N GPUs. usage: python -m torch.distributed.run --nproc-per-node N profiles/d2h_concurrent.py"""
import json, os, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.copy_(d, non_blocking=True); torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(12): h.copy_(d, non_blocking=True)
e1.record(); torch.cuda.synchronize()
gbs = torch.tensor([12 * n / (e0.elapsed_time(e1) / 1e3) / 1e9], device="cuda", dtype=torch.float64)
allg = [torch.zeros_like(gbs) for _ in range(world)]
if world > 1: dist.all_gather(allg, gbs)
else: allg = [gbs]
if rank == 0:
    per = [float(x) for x in allg]
    os.write(1, (json.dumps({"n_gpus": world, "per_gpu_GBps": [round(p, 1) for p in per], "aggregate_GBps": round(sum(per), 1)}) + "\n").encode())
if world > 1: dist.destroy_process_group()
