"""Host-link probe under torchrun (one rank per GPU): NUMA node and local CPUs of every GPU, the CPUs / memory nodes the
process may use, H2D bandwidth of a 256 MiB pinned buffer alone and with all ranks busy, allocated as-is and after
pinning the thread to the GPU's local CPUs.  Prints one line per rank."""
import os, time, json
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
def barrier():
    if world > 1: dist.barrier()
p = torch.cuda.get_device_properties(lr)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
def rd(path):
    try: return open(path).read().strip()
    except Exception as e: return "?"
info = {"rank": rank, "bus": bus, "gpu_numa": rd(f"/sys/bus/pci/devices/{bus}/numa_node"), "gpu_cpus": rd(f"/sys/bus/pci/devices/{bus}/local_cpulist"),
        "allowed_cpus": len(os.sched_getaffinity(0)), "cpu_min_max": (min(os.sched_getaffinity(0)), max(os.sched_getaffinity(0))),
        "nodes_online": rd("/sys/devices/system/node/online"),
        "mems_allowed": [l.split(":")[1].strip() for l in open("/proc/self/status") if l.startswith("Mems_allowed_list")][0]}
N = 256 << 20
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
def bw(host, concurrent):
    best = 0.0
    for rep in range(3):
        if concurrent: barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6): dev.copy_(host, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        best = max(best, 6 * N / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return round(best, 1)
def parse(cl):
    out = set()
    for part in cl.split(","):
        if "-" in part: a, b = part.split("-"); out |= set(range(int(a), int(b) + 1))
        elif part.strip().isdigit(): out.add(int(part))
    return out
h0 = torch.empty(N, dtype=torch.uint8).pin_memory(); h0.fill_(1)
for r in range(world):          # alone, one rank at a time
    barrier()
    if r == rank: info["alone_default"] = bw(h0, False)
barrier(); info["busy_default"] = bw(h0, True)
local = parse(info["gpu_cpus"]) & os.sched_getaffinity(0)
info["local_allowed"] = len(local)
if local:
    old = os.sched_getaffinity(0); os.sched_setaffinity(0, local)
    h1 = torch.empty(N, dtype=torch.uint8).pin_memory(); h1.fill_(1)
    for r in range(world):
        barrier()
        if r == rank: info["alone_local"] = bw(h1, False)
    barrier(); info["busy_local"] = bw(h1, True)
    os.sched_setaffinity(0, old)
else:
    for r in range(world): barrier()
    barrier(); barrier(); barrier(); barrier()
# device -> host too
hd = torch.empty(N, dtype=torch.uint8).pin_memory()
barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(6): hd.copy_(dev, non_blocking=True)
e1.record(); torch.cuda.synchronize()
info["busy_d2h"] = round(6 * N / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
print(json.dumps(info), flush=True)
if world > 1: dist.destroy_process_group()
