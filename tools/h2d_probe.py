"""Host<->device copy bandwidth with every rank copying at once (torchrun, one rank per GPU).

Answers what the sharded end-to-end path can reach on a box: per-rank H2D / D2H / both-way
bandwidth alone and with all ranks active, with pinned buffers placed by default and on the cores
next to the GPU.  Prints one JSON line (rank 0).
"""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gpu_local_cpus  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # pragma: no cover
        return repr(e)


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gib = float(os.environ.get("PROBE_GIB", "2"))
    n = int(gib * (1 << 30)) // 8
    d_a = torch.empty(n, dtype=torch.complex64, device=dev)
    d_b = torch.empty(n, dtype=torch.complex64, device=dev)
    s2 = torch.cuda.Stream(dev)
    out = {"world": world, "gib_per_copy": gib}
    if rank == 0:
        out["topo"] = sh("nvidia-smi topo -m | head -14")
        out["numa"] = sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'")
        out["affinity"] = len(os.sched_getaffinity(0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, active=True, reps=3):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if active:
            for _ in range(reps):
                fn()
        e1.record()
        s2.synchronize()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3 if active else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return reps * gib * 1.073741824 / float(t.item())        # GB/s per rank (slowest rank)

    for place in ("default", "local"):
        if place == "local":
            with gpu_local_cpus(local) as numa:
                h_a = torch.empty(n, dtype=torch.complex64).pin_memory()
                h_b = torch.empty(n, dtype=torch.complex64).pin_memory()
                h_a.zero_(); h_b.zero_()
            out["local_info_rank0"] = numa.info
        else:
            h_a = torch.empty(n, dtype=torch.complex64).pin_memory()
            h_b = torch.empty(n, dtype=torch.complex64).pin_memory()
            h_a.zero_(); h_b.zero_()

        def h2d():
            d_a.copy_(h_a, non_blocking=True)

        def d2h():
            h_b.copy_(d_b, non_blocking=True)

        def both():
            d_a.copy_(h_a, non_blocking=True)
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)

        def both_wait():
            both()
            torch.cuda.current_stream().wait_stream(s2)
        h2d(); d2h(); barrier()
        res = {}
        res["h2d_rank0_alone"] = timed(h2d, active=(rank == 0))
        res["h2d_all"] = timed(h2d)
        res["d2h_all"] = timed(d2h)
        res["both_all_per_direction"] = timed(both_wait)
        if world >= 4:
            res["h2d_even_ranks"] = timed(h2d, active=(rank % 2 == 0)) 
            res["h2d_first_half"] = timed(h2d, active=(rank < world // 2))
        out[place] = {k: round(v, 1) for k, v in res.items()}
        del h_a, h_b
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
