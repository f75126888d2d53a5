"""Host-link ceiling probe (NOT product path; nvbandwidth-style): every rank copies 1 GiB pinned host buffers to its GPU, from its
GPU, and both ways at once, all ranks at the same time.  The aggregate GB/s per direction at N = 1, 2, 4, 8 is the ceiling the
end-to-end arm of bench.py (4.3 GB up + 5.4 GB down per rank and step) runs against.
    python tools/probe_host_link.py                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/probe_host_link.py
Prints one JSON line on rank 0 (and appends it to gpurun_out/host_link.jsonl)."""
import json, os, time
import torch

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
up = torch.empty(n, dtype=torch.uint8).pin_memory(); down = torch.empty(n, dtype=torch.uint8).pin_memory()
up.fill_(1)
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, k=6):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                a.copy_(up, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                down.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return k * n / float(t.item()) / 1e9          # GB/s per rank and direction, slowest rank


out = {"world": world, "bytes_per_copy": n, "cpus_allowed": len(os.sched_getaffinity(0))}
for mode in ("h2d", "d2h", "both"):
    run(mode, 2)
    per_rank = max(run(mode) for _ in range(2))
    out[mode] = {"per_rank_gbs_each_direction": round(per_rank, 1), "aggregate_gbs_each_direction": round(per_rank * world, 1)}
if rank == 0:
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/host_link.jsonl", "a").write(json.dumps(out) + "\n")
if dist is not None:
    dist.destroy_process_group()
