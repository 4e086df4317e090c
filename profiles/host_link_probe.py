#!/usr/bin/env python
"""Host<->device link probe for the e2e arm: NUMA placement of GPU 0, CPU affinity, and pinned H2D / D2H bandwidth
with the process unbound and bound to the GPU's NUMA node.  python profiles/host_link_probe.py"""
import os
import time

import torch


def cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if "-" in part:
            a, b = part.split("-"); out |= set(range(int(a), int(b) + 1))
        elif part:
            out.add(int(part))
    return out


def bw(n_bytes=100 << 20, reps=10):
    h = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda:0")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    h2d = n_bytes * reps / (time.perf_counter() - t) / 1e9
    t = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    d2h = n_bytes * reps / (time.perf_counter() - t) / 1e9
    return h2d, d2h


def main():
    p = torch.cuda.get_device_properties(0)
    bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    node = -1
    try:
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
    except OSError as e:
        print("numa_node unreadable:", e)
    aff = os.sched_getaffinity(0)
    print(f"gpu0 {p.name} pci {bdf} numa_node {node}; affinity {len(aff)} cpus; nodes:",
          sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")))
    print("unbound  H2D %.1f GB/s  D2H %.1f GB/s" % bw())
    if node >= 0:
        cpus = cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & aff
        if cpus:
            os.sched_setaffinity(0, cpus)
            print(f"bound to node {node} ({len(cpus)} cpus)  H2D %.1f GB/s  D2H %.1f GB/s" % bw())
    os.system("nvidia-smi topo -m 2>/dev/null | head -12")
    os.system("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv 2>/dev/null")


if __name__ == "__main__":
    main()
