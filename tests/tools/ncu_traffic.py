"""DRAM traffic per step of the kernels whose name matches a regex, from an `ncu --csv --metrics
dram__bytes_read.sum,dram__bytes_write.sum` log:  python tests/tools/ncu_traffic.py log.csv REGEX STEP_KERNEL
(STEP_KERNEL: a kernel launched once per step, to count the steps).  Prints one JSON object."""
import csv
import json
import re
import sys


def main():
    path, rx, step_kernel = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
    h = rows[hdr]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per, launches, ids = {}, {}, set()
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[mi].startswith("dram__bytes"):
            continue
        name = r[ki]
        if not rx.search(name):
            continue
        per[name] = per.get(name, 0.0) + float(r[vi].replace(",", "")) * mult.get(r[ui], 1.0)
        if r[mi] == "dram__bytes_read.sum":
            launches[name] = launches.get(name, 0) + 1
    steps = max(1, sum(v for k, v in launches.items() if step_kernel in k))
    print(json.dumps({"steps": steps, "bytes_per_step": sum(per.values()) / steps,
                      "per_kernel_per_step": {k[:60]: v / steps for k, v in per.items()},
                      "launches_per_step": {k[:60]: v / steps for k, v in launches.items()}}))


if __name__ == "__main__":
    main()
