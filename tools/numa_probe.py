"""What the box exposes about GPU <-> NUMA placement (development probe for bench.py's bind_to_gpu_numa_node)."""
import glob
import os

import pynvml

pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("cpus allowed:", len(os.sched_getaffinity(0)), "nodes:", sorted(glob.glob("/sys/devices/system/node/node*")))
for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    print(p, open(p).read().strip())
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    sysfs = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
    node = open(sysfs).read().strip() if os.path.exists(sysfs) else "missing"
    try:
        aff = list(pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64))
    except Exception as e:
        aff = repr(e)
    try:
        numa = pynvml.nvmlDeviceGetNumaNodeId(h)
    except Exception as e:
        numa = repr(e)[:60]
    print(i, bus, "sysfs numa_node", node, "nvml cpu affinity", [hex(a) for a in aff] if isinstance(aff, list) else aff, "nvml numa id", numa)
