#!/usr/bin/env python3
"""Experiment support: concurrent D2H bandwidth of every visible GPU with pinned host buffers, with and
without binding each worker to the CPUs local to its GPU (sysfs numa_node / local_cpulist).
    python tools/pcie_probe.py [--bind]"""
import os
import sys
import time

import torch
import torch.multiprocessing as mp


def local_cpus(dev: int):
    bus = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
    except Exception as exc:  # noqa: BLE001
        return None, f"nvml: {exc}"
    bus = bus.lower()
    if len(bus.split(":")[0]) == 8:
        bus = bus[4:]
    path = f"/sys/bus/pci/devices/{bus}"
    try:
        node = open(path + "/numa_node").read().strip()
        cpus = open(path + "/local_cpulist").read().strip()
    except OSError as exc:
        return None, f"sysfs: {exc}"
    out = []
    for part in cpus.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out, f"bus {bus} numa_node {node} cpus {cpus}"


def worker(dev, bind, barrier, q):
    info = ""
    if bind:
        cpus, info = local_cpus(dev)
        allowed = os.sched_getaffinity(0)
        if cpus:
            use = [c for c in cpus if c in allowed] or list(allowed)
            os.sched_setaffinity(0, use)
    torch.cuda.set_device(dev)
    n = 1 << 28
    d = torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.copy_(d, non_blocking=True)
    bidir = "--bidir" in sys.argv
    d2 = torch.empty(n // 4, dtype=torch.uint8, device=f"cuda:{dev}")
    h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
    s2 = torch.cuda.Stream()
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.time()
    for _ in range(8):
        h.copy_(d, non_blocking=True)
        if bidir:
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    q.put((dev, 8 * n / dt / 1e9, info))


if __name__ == "__main__":
    bind = "--bind" in sys.argv
    mp.set_start_method("spawn")
    g = torch.cuda.device_count()
    print("affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus; gpus", g)
    for group in ([0], list(range(g))):
        barrier = mp.Barrier(len(group))
        q = mp.Queue()
        ps = [mp.Process(target=worker, args=(d, bind, barrier, q)) for d in group]
        [p.start() for p in ps]
        res = sorted(q.get() for _ in ps)
        [p.join() for p in ps]
        print("bind" if bind else "nobind", "gpus", group, [(d, round(bw, 1)) for d, bw, _ in res], "total", round(sum(bw for _, bw, _ in res), 1), "GB/s")
        if bind:
            print("  ", [i for _, _, i in res])
