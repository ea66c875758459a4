"""Raw host-link ceiling: concurrent pinned D2H copies from every rank of the job, no solve.

    python scripts/pcie_probe.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/pcie_probe.py                                     # N GPUs of one box, all copying at once

For each host-buffer kind (cudaHostAlloc = torch pin_memory; mmap + MADV_HUGEPAGE + cudaHostRegister =
dynode_b200.hostmem) every rank copies `--gib` GiB device -> host `--reps` times in `--chunk-mb` pieces on one stream,
bracketed by a barrier; the aggregate is all ranks' bytes over the slowest rank's time.  Prints one JSON line per kind
on rank 0 (kept under profiles/): this is the denominator `bench.py` reports as e2e.pcie_ceiling_gbs.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--chunk-mb", type=int, default=256)
    ap.add_argument("--affinity", action="store_true", help="pin each rank to its own slice of the host cores")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.affinity:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        info = {"world": world, "host_cores": os.cpu_count(),
                "thp": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
                "numa": subprocess.run("lscpu | grep -i numa", shell=True, capture_output=True, text=True).stdout.split("\n"),
                "affinity": args.affinity}
        print(json.dumps(info), flush=True)
        if world > 1:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)

    from dynode_b200 import hostmem
    n = int(args.gib * (1 << 30)) // 8
    chunk = args.chunk_mb * (1 << 20) // 8
    d = torch.empty(n, dtype=torch.float64, device=dev).normal_()
    for kind in ("cudaHostAlloc", "thp+cudaHostRegister", "cudaHostAlloc_again"):
        t_alloc = time.perf_counter()
        if kind.startswith("cudaHostAlloc"):
            h = torch.empty(n, dtype=torch.float64).pin_memory()
            huge = None
        else:
            h = hostmem.pinned_empty((n,))
            huge = h._dynode_host_buffer.huge_bytes()
        t_alloc = time.perf_counter() - t_alloc
        res = {}
        for direction in ("d2h", "h2d"):
            src, dst = (d, h) if direction == "d2h" else (h, d)
            for lo in range(0, n, chunk):
                dst[lo:lo + chunk].copy_(src[lo:lo + chunk], non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                for lo in range(0, n, chunk):
                    dst[lo:lo + chunk].copy_(src[lo:lo + chunk], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[direction] = {"rank0_gbs": args.reps * n * 8 / dt / 1e9,
                              "aggregate_gbs": world * args.reps * n * 8 / float(t.item()) / 1e9}
        if rank == 0:
            print(json.dumps({"kind": kind, "world": world, "gib_per_rank": args.gib, "chunk_mb": args.chunk_mb,
                              "alloc_s": t_alloc, "huge_bytes": huge, **res}), flush=True)
        del h
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
