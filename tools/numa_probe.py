"""H2D bandwidth of an 84 MiB pinned buffer with the process bound to each NUMA node in turn (first touch places the pages)."""
import glob, os, subprocess, time
import torch

def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
print("numa nodes:", {k: (v[0], v[-1], len(v)) for k, v in nodes.items() if v})
print("initial affinity:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...")
torch.cuda.init()
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
bus = pynvml.nvmlDeviceGetPciInfo(h).busId
bus = bus.decode() if isinstance(bus, bytes) else bus
p = "/sys/bus/pci/devices/" + bus[-12:].lower() + "/numa_node"
print("gpu0 bus", bus, "numa_node", open(p).read().strip() if os.path.exists(p) else "?")
allowed = os.sched_getaffinity(0)
d = torch.empty(84 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def measure(tag):
    a = torch.empty(84 * 1024 * 1024, dtype=torch.uint8)
    a.fill_(1)
    a = a.pin_memory()
    for _ in range(3):
        d.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        d.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{tag}: H2D 84 MiB {dt * 1e3:.3f} ms = {a.numel() / dt / 1e9:.1f} GB/s")

measure("unbound")
for k, cpus in nodes.items():
    use = set(cpus) & allowed
    if not use:
        print("node", k, "not allowed"); continue
    os.sched_setaffinity(0, use)
    measure(f"bound to node {k} ({len(use)} cpus)")
os.sched_setaffinity(0, allowed)
