#!/usr/bin/env python
"""Host-side diagnostics of the upload path: NUMA visibility (sysfs / NVML), CPU affinity, pinned H2D rate of one band.
usage: python tools/host_topology.py [--mb 128]"""
import argparse
import glob
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mb', type=int, default=128)
    a = ap.parse_args()
    import torch
    n = torch.cuda.device_count()
    print("gpus", n, "cpus allowed", len(os.sched_getaffinity(0)), "of", os.cpu_count())
    print("numa nodes (sysfs):", sorted(glob.glob('/sys/devices/system/node/node*')))
    for i in range(n):
        pr = torch.cuda.get_device_properties(i)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        p = "/sys/bus/pci/devices/%s/numa_node" % bus
        try:
            node = open(p).read().strip()
        except Exception as e:
            node = "unreadable (%s)" % type(e).__name__
        print("gpu", i, bus, "sysfs numa_node:", node)
    try:
        import pynvml
        pynvml.nvmlInit()
        for i in range(n):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            try:
                aff = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
                cpus = [64 * w + b for w, v in enumerate(aff) for b in range(64) if (v >> b) & 1]
                print("gpu", i, "nvml cpu affinity: %d cpus, first %s last %s" % (len(cpus), cpus[:1], cpus[-1:]))
            except Exception as e:
                print("gpu", i, "nvml cpu affinity failed:", e)
            try:
                maff = pynvml.nvmlDeviceGetMemoryAffinity(h, 4, pynvml.NVML_AFFINITY_SCOPE_NODE)
                print("gpu", i, "nvml memory affinity (node mask words):", list(maff))
            except Exception as e:
                print("gpu", i, "nvml memory affinity failed:", e)
    except Exception as e:
        print("pynvml unavailable:", e)
    dev = torch.device('cuda:0')
    nb = a.mb << 20
    host = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    d = torch.empty(nb, dtype=torch.uint8, device=dev)
    for piece in (nb // 16, nb // 4, nb, nb // 2, nb // 8, nb // 32, nb, nb // 16):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in range(4):
            for o in range(0, nb, piece):
                d[o:o + piece].copy_(host[o:o + piece], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print("H2D %d MB in pieces of %d MB: %.2f ms = %.1f GB/s" % (a.mb, piece >> 20, dt * 1e3, nb / dt / 1e9))


if __name__ == '__main__':
    main()
