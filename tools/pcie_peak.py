"""Plain pinned-memory copy bandwidth of the box on N GPUs at once (H2D alone, D2H alone, both directions together): the
ceiling of bench.py's end-to-end leg.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_peak.py > profiles/pcie_ceiling_Ngpu.json

Every rank copies at the same time (barrier, then timed copies); rank 0 prints ONE JSON object with the per-rank and the
aggregate rates and the time the box needs for the bytes of one end-to-end bench step (189 MB up + 94 MB down per GPU).
argument `numa`: bind the process to the CPUs of the GPU's NUMA node (from sysfs) before the pinned buffers are allocated."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def bind_numa(local):
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        import subprocess
        q = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        node = open(f"/sys/bus/pci/devices/{q.lower()[4:]}/numa_node").read().strip()
        cpus = open(f"/sys/devices/system/node/node{max(int(node), 0)}/cpulist").read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return dict(node=int(node), cpus=cpus)
    except Exception as e:  # noqa: BLE001
        return dict(error=repr(e))


def main():
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    numa = bind_numa(local) if "numa" in sys.argv[1:] else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = 512 * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n // 2, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(f, reps=6):
        f()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    a = timed(lambda: d.copy_(h, non_blocking=True))
    b = timed(lambda: h2.copy_(d2, non_blocking=True))
    c = timed(both)
    mine = torch.tensor([n / a / 1e9, n / 2 / b / 1e9, n / c / 1e9, n / 2 / c / 1e9, c * 189136896 / n * 1e3], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        rows = [[float(v) for v in r.tolist()] for r in allr]
        out = {
            "n_gpus": world, "numa_binding": numa, "host_cores": len(os.sched_getaffinity(0)) if numa is None else os.cpu_count(),
            "per_rank_gbs": [dict(h2d_alone=r[0], d2h_alone=r[1], h2d_duplex=r[2], d2h_duplex=r[3]) for r in rows],
            "aggregate_gbs": dict(h2d_alone=sum(r[0] for r in rows), d2h_alone=sum(r[1] for r in rows),
                                  duplex_total=sum(r[2] + r[3] for r in rows)),
            "e2e_step_bytes_ms": max(r[4] for r in rows),
            "note": "all ranks copy at the same time; e2e_step_bytes_ms = time for 189 MB up + 94 MB down per GPU (one bench.py e2e step of configs[1]) as plain pinned copies, max over ranks",
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
