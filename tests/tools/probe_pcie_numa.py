"""Host<->device copy bandwidth of pinned buffers allocated from each NUMA node of the box (one GPU).

Prints one JSON object: the GPU's PCI address and NUMA node as sysfs reports them, the CPU list of every node, and
the measured H2D / D2H GB/s of a 256 MB pinned buffer allocated while the calling thread was bound to that node.
Used to decide where the context's staging buffers have to live (hfb_ctx_create binds the allocating thread).
"""
import glob
import json
import os
import sys

import torch


def cpulist(txt):
    out = []
    for part in txt.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def main():
    dev = torch.device("cuda:0")
    torch.cuda.init()
    prop = torch.cuda.get_device_properties(0)
    bdf = None
    try:
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
    except AttributeError:
        pass
    gpu_node = None
    if bdf:
        p = "/sys/bus/pci/devices/%s/numa_node" % bdf
        if os.path.exists(p):
            gpu_node = int(open(p).read())
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
    allowed = sorted(os.sched_getaffinity(0))
    res = {"gpu_bdf": bdf, "gpu_numa_node": gpu_node, "allowed_cpus": len(allowed), "nodes": {}}
    n = 256 << 20
    g = torch.empty(n, dtype=torch.uint8, device=dev)
    for k, (node, cpus) in enumerate(sorted(nodes.items())):
        use = sorted(set(cpus) & set(allowed))
        if not use:
            res["nodes"][node] = {"cpus": len(cpus), "usable": 0}
            continue
        os.sched_setaffinity(0, use)
        h = torch.empty(n + 4096 * (k + 1), dtype=torch.uint8).pin_memory()  # a fresh cudaHostAlloc every time
        h[:n].fill_(1)
        out = {"cpus": len(cpus), "usable": len(use)}
        for name, (dst, src) in {"h2d": (g, h[:n]), "d2h": (h[:n], g)}.items():
            best = 0.0
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dst.copy_(src, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
            out[name + "_GBps"] = round(best, 2)
        # both directions at once (what a pipelined batch call does)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        g2 = torch.empty(n, dtype=torch.uint8, device=dev)
        h2 = torch.empty(n + 8192 * (k + 1), dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            g.copy_(h[:n], non_blocking=True)
        with torch.cuda.stream(s2):
            h2[:n].copy_(g2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        out["both_GBps_each"] = round(n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
        res["nodes"][node] = out
        os.sched_setaffinity(0, allowed)
    print(json.dumps(res))


if __name__ == "__main__":
    sys.exit(main())
