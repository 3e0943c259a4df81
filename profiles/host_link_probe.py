"""What one host can move to and from its GPUs at once: every rank copies the e2e leg's bytes (2.0 GB pinned -> device, 1.13 GB device ->
pinned) with plain cudaMemcpyAsync on two streams, no kernels; ranks 0..k-1 take part for k = 1, 2, 4, 8. Launch with torchrun, one rank per
GPU. Prints one JSON object from rank 0: seconds per round and aggregate GB/s per k."""
import json, os, time
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
IN, OUT = 2_002_400_000, 1_128_800_000
h_in = torch.empty(IN, dtype=torch.uint8, pin_memory=True); d_in = torch.empty(IN, dtype=torch.uint8, device="cuda")
h_out = torch.empty(OUT, dtype=torch.uint8, pin_memory=True); d_out = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
out = {}
for k in [x for x in (1, 2, 4, 8) if x <= world]:
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank < k:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
            s1.synchronize(); s2.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep:
            best = min(best, float(dt.item()))
    out[f"{k}_ranks"] = {"seconds": best, "aggregate_GBps": k * (IN + OUT) / best / 1e9, "per_rank_GBps": (IN + OUT) / best / 1e9}
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
