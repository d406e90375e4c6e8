"""Host<->device copy rates with N ranks copying AT THE SAME TIME (one process per GPU under torchrun): the table VERDICT r01
item 4 asked for before touching the e2e path. Every rank copies 1 GiB pinned<->device three times per direction.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ranks_probe.py
"""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 28
    h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
    d_a, d_b = torch.empty(n, dtype=torch.float32, device="cuda"), torch.ones(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    out = {}
    for name, fn, gb in (("h2d", h2d, 4 * n / 1e9), ("d2h", d2h, 4 * n / 1e9), ("duplex", both, 8 * n / 1e9)):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        t = torch.tensor([gb / dt], device="cuda", dtype=torch.float64)
        if world > 1:
            lo, tot = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            out[name] = (float(lo.item()), float(tot.item()))
        else:
            out[name] = (float(t.item()), float(t.item()))
    if rank == 0:
        print(f"ranks {world}: " + "  ".join(f"{k} min/rank {v[0]:.1f} GB/s, aggregate {v[1]:.1f} GB/s" for k, v in out.items()), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
