#!/usr/bin/env python
"""bench.py — throughput of the chunk-meshing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step = one pass of the mesher over one batch of 4096 chunks (64 KiB each, 256 MiB per batch).  Four distinct
batches (seeds 42..45, 1 GiB in total, > 126 MB of L2) rotate so no step re-reads what the previous one left in
L2.  Prints ONE JSON line (rank 0).

workloads (BASELINE.json configs):
  terrain_greedy   configs[2]  4096 noise-terrain chunks [0,16)^3, greedy quads   (default: the north-star target mode)
  terrain_culled   configs[1]  same batches, culled faces
  checker_culled / checker_greedy   configs[3]  world-parity checkerboard, 98 304 quads per chunk
  world_fused      configs[4]  65^3 = 274 625 chunk coords, fused terrain+greedy, sharded (strong scaling)

`--impl reference`: the reference snapshot has no implementation of this path, so that arm times the
in-repo CPU oracle (kind "port") with all host threads on the SAME batch, coordinates and neighbour table.  It never
imports the product package: libvxp.so is not mapped in that process.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CHUNKS = 4096
N_BATCHES = 4
CHUNK_BYTES = 65536
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent
WORLD_LO, WORLD_HI = (-32, -32, -32), (33, 33, 33)   # config 5: 65^3 chunk coordinates centred on the origin


# --------------------------------------------------------------------------- helpers (numpy only: shared by both arms)
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def grid_coords_np(lo, hi) -> np.ndarray:
    """Chunk coordinates of the box [lo, hi), x fastest (the order of voxelphile_b200.grid_coords)."""
    z, y, x = np.meshgrid(np.arange(lo[2], hi[2]), np.arange(lo[1], hi[1]), np.arange(lo[0], hi[0]), indexing="ij")
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int32)


def neighbour_table_np(lo, hi) -> np.ndarray:
    """nbr6 of a full box of coordinates in grid_coords order, pure numpy (SPEC §1): -1 outside the box."""
    dims = np.array(hi) - np.array(lo)
    nx, ny, nz = (int(d) for d in dims)
    idx = np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)
    nbr = np.full((nz, ny, nx, 6), -1, dtype=np.int32)
    nbr[:, :, 1:, 0] = idx[:, :, :-1]
    nbr[:, :, :-1, 1] = idx[:, :, 1:]
    nbr[:, 1:, :, 2] = idx[:, :-1, :]
    nbr[:, :-1, :, 3] = idx[:, 1:, :]
    nbr[1:, :, :, 4] = idx[:-1, :, :]
    nbr[:-1, :, :, 5] = idx[1:, :, :]
    return np.ascontiguousarray(nbr.reshape(-1, 6))


def algorithmic_bytes(counts: np.ndarray) -> int:
    """SPEC §8: read every voxel once, write 56 B per quad and the 8-byte meta entry."""
    return int(counts.shape[0]) * (CHUNK_BYTES + 8) + 56 * int(counts.sum())


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py traffic from an `ncu --set full` report)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t["kernels"].get(kernel)
        return (e["dram_read_bytes"] + e["dram_write_bytes"], {"file": "profiles/ncu_traffic.json", "commit": t.get("commit"),
                                                                 "report": e.get("report")}) if e else (None, None)
    except Exception:
        return None, None


def mix64_np(keys: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = keys.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def chunk_hashes_np(verts: np.ndarray, meta: np.ndarray, total: int) -> np.ndarray:
    """Order-independent per-chunk hash (SPEC §4) of a host mesh: keys from vertex 0 of every quad (SPEC §5 inverse)."""
    out = np.zeros(meta.shape[0], dtype=np.uint64)
    if total == 0:
        return out
    v0 = verts.reshape(-1, 4)[:total, 0]
    w0, w1 = (v0 & np.uint64(0xFFFFFFFF)).astype(np.int64), (v0 >> np.uint64(32)).astype(np.int64)
    x, y, z, f = w0 & 63, (w0 >> 6) & 63, (w0 >> 12) & 63, (w0 >> 18) & 7
    axis = f >> 1
    along = np.where(axis == 0, x, np.where(axis == 1, y, z)) - (f & 1)
    u = np.where(axis == 0, y, x)
    v = np.where(axis == 2, y, z)
    key = u | v << 6 | along << 12 | f << 18 | ((w1 >> 16) & 31) << 21 | ((w1 >> 21) & 31) << 26 | (w1 & 0xFFFF) << 31
    h = mix64_np(key.astype(np.uint64))
    owner = np.zeros(total, dtype=np.int64)
    first, cnt = meta[:, 0].astype(np.int64), meta[:, 1].astype(np.int64)
    for i in np.nonzero(cnt)[0]:
        owner[first[i]:first[i] + cnt[i]] = i
    with np.errstate(over="ignore"):
        np.add.at(out, owner, h)
    return out


class ClockSampler:
    """SM clock and throttle reasons sampled (NVML, in-process, every 50 ms) while the GPU is under load.
    NVML is initialised before the warm-up so its start-up cost never lands in the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.t, self.stop_flag, self.h, self.nv = index, [], None, False, None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            try:
                self.h = nv.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid) if uuid else None
            except Exception:
                self.h = None
            if self.h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
                self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.nv = nv
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            # the first call of each NVML query takes milliseconds (measured: 11.6 ms / 4.4 ms, later ones
            # ~1 us) and holds driver locks the launch path needs: pay that here, not in the timed region
            for _ in range(2):
                nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                self._reasons()
                nv.nvmlDeviceGetUtilizationRates(self.h)
        except Exception:
            self.nv = None

    def _reasons(self):
        nv = self.nv
        if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons"):
            return nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        return nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = self._reasons()
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.rows.append((time.time(), float(mhz), int(reasons), int(util)))
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self, t0: float, t1: float, window: str):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "window": "NVML unavailable"}
        self.stop_flag = True
        self.t.join(timeout=1)
        nv = self.nv
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        if not rows:
            rows, window = [r for r in self.rows if r[3] > 0] or self.rows, "whole run under load (timed region shorter than one sample)"
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = [n for n, b in bits.items() if any(r[2] & b for r in rows)]
        sm = [r[1] for r in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(rows), "window": window}


def workload_box(workload: str):
    return (WORLD_LO, WORLD_HI) if workload == "world_fused" else ((0, 0, 0), (16, 16, 16))


# --------------------------------------------------------------------------- reference arm (CPU oracle)
def run_reference(args):
    """The reference has no implementation of this path; the comparand is the in-repo CPU oracle
    (kind "port"), all host threads, on the GPU arm's own configuration: the same 4096 coordinates, the same
    neighbour table, batches of seeds 42 and 43 alternating (world_fused: a bounded slab of the same world)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    mode = 0 if args.workload.endswith("culled") else 1
    threads = O.max_threads()
    lo, hi = workload_box(args.workload)
    same_config = True
    if args.workload == "world_fused":
        # bounded sample of the same world: two z-layers (8450 chunks) per step through the oracle's world mesher,
        # which generates every chunk and neighbour itself (no voxels stored: what the fused GPU path replaces)
        slab_lo, slab_hi = (lo[0], lo[1], -1), (hi[0], hi[1], 1)
        n_step = 65 * 65 * 2
        sample = "two z-layers (z = -1, 0: 8450 chunks) of the 65^3 world per step, terrain generated and meshed by the oracle"
        same_config = False

        def step(s):
            O.terrain_world_hash(42, slab_lo, slab_hi, mode, threads)
    elif args.workload.startswith("checker"):
        coords = grid_coords_np(lo, hi)
        nbr = neighbour_table_np(lo, hi)
        vox = np.stack([O.pattern_fill(0, *map(int, c)) for c in coords])
        n_step = coords.shape[0]
        cap = int(O.mesh_batch_hash(vox, nbr, mode, threads)[1].sum())
        sample = "the 4096-chunk checkerboard batch ([0,16)^3), every step"

        def step(s):
            O.mesh_batch(vox, nbr, mode, threads, capacity=cap)
    else:
        coords = grid_coords_np(lo, hi)
        nbr = neighbour_table_np(lo, hi)
        batches = [O.terrain_fill_batch(42 + k, coords, threads) for k in range(2)]
        n_step = coords.shape[0]
        cap = int(max(O.mesh_batch_hash(v, nbr, mode, threads)[1].sum() for v in batches))
        sample = "the GPU arm's 4096-chunk batches of seeds 42 and 43 alternating (same coordinates, same neighbour table), every step"

        def step(s):
            O.mesh_batch(batches[s % 2], nbr, mode, threads, capacity=cap)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        step(s)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    value = n_step / per_step
    line = {
        "impl": "reference", "metric": "chunks meshed/sec", "value": value, "unit": "chunks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": args.workload, "chunks_per_step": int(n_step), "mode": "greedy" if mode else "culled",
                   "same_config": same_config,
                   "note": "the reference snapshot has no mesher: this is the in-repo CPU oracle (SPEC v1), not voxelphile"},
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", default="terrain_greedy",
                    choices=["terrain_culled", "terrain_greedy", "checker_culled", "checker_greedy", "world_fused"])
    ap.add_argument("--impl", default="vxp", choices=["vxp", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-also", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--soak-seconds", type=float, default=0.3, help="minimum warm-up duration (0 under ncu)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-overlap", action="store_true",
                    help="serialise consecutive launches (vxp_ctx_set_overlap off) in the timed region")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from voxelphile_b200 import CULLED, GREEDY, PATTERN_CHECKER, Mesher
    from voxelphile_b200.shard import gather_counts, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    peak, peak_src = hbm_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_ok(flag: bool) -> bool:
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # one explicit non-default stream for everything: torch allocations/fills, our kernels and the
    # timing events all live on it (torch.cuda.Event only sees torch's current stream; the Mesher follows it)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    m = Mesher(local)
    mode = CULLED if args.workload.endswith("culled") else GREEDY
    fused = args.workload == "world_fused"
    lo_box, hi_box = workload_box(args.workload)
    coords_np = grid_coords_np(lo_box, hi_box)

    # ---------------- inputs resident in HBM
    if fused:
        lo, hi = partition(coords_np.shape[0], rank, world)       # strong scaling: shard the world by coordinate ranges
        my_coords = torch.from_numpy(np.ascontiguousarray(coords_np[lo:hi])).to(dev)
        n_step = hi - lo
        batches = [None]
        scaling = "strong"
        nbr_np = dnbr = dcoords = None
    else:
        nbr_np = neighbour_table_np(lo_box, hi_box)
        dcoords = torch.from_numpy(coords_np).to(dev)
        dnbr = torch.from_numpy(nbr_np).to(dev)
        n_step = N_CHUNKS
        batches = []
        seeds = []
        for k in range(N_BATCHES):
            if args.workload.startswith("checker"):
                shift = torch.tensor([[16 * k, 0, 0]], dtype=torch.int32, device=dev)
                batches.append(m.pattern_fill((dcoords + shift).contiguous(), PATTERN_CHECKER))
                seeds.append(None)
            else:
                seeds.append(42 + N_BATCHES * rank + k)                   # weak scaling: own seeds per rank
                batches.append(m.terrain_fill(dcoords, seeds[-1]))
        scaling = "weak"
    torch.cuda.synchronize()

    def launch(k, out):
        if fused:
            return m.terrain_mesh(my_coords, 42, mode, out=out)
        return m.mesh(batches[k % len(batches)], dnbr, mode, out=out)

    # ---------------- size the output arena with one untimed pass per batch
    probe = m.alloc_mesh(n_step, 0)
    totals, counts = [], []
    for k in range(len(batches)):
        launch(k, probe)
        torch.cuda.synchronize()
        totals.append(int(probe.total.item()))
        counts.append(probe.meta[:, 1].cpu().numpy().astype(np.int64))
    # two output meshes used alternately: launch k+1 may start while launch k drains (vxp_ctx_set_overlap), and a
    # consumer would read mesh k while mesh k+1 is produced
    outs = [m.alloc_mesh(n_step, max(totals) + 1), m.alloc_mesh(n_step, max(totals) + 1)]
    out = outs[0]
    overlap = not args.no_overlap and not fused
    m.set_overlap(overlap)
    alg_bytes = [algorithmic_bytes(c) if not fused else 12 * n_step + 56 * int(c.sum()) + 8 * n_step for c in counts]

    # ---------------- parity check of this rank's shard against the CPU oracle (outside the timed region): per-chunk
    # quad counts and key hashes of one batch (fused: of a bounded slab of this rank's range)
    parity = None
    if not args.no_cpu and not args.workload.startswith("checker"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        launch(0, outs[0])
        host = outs[0].to_host()
        got_h = chunk_hashes_np(host.verts, host.meta, host.total_quads)
        if fused:
            # whole z-layers inside [lo, hi): the first full layer of the rank's range (4225 chunks)
            layer = 65 * 65
            z0 = -(-lo // layer)
            if (z0 + 1) * layer <= hi:
                zc = WORLD_LO[2] + z0
                want_h, want_c = O.terrain_world_hash(42, (WORLD_LO[0], WORLD_LO[1], zc), (WORLD_HI[0], WORLD_HI[1], zc + 1), 1)
                a = z0 * layer - lo
                ok = np.array_equal(host.meta[a:a + layer, 1], want_c) and np.array_equal(got_h[a:a + layer], want_h)
                checked = layer
            else:
                ok, checked = True, 0
        else:
            vox0 = batches[0].cpu().numpy().view(np.uint16)
            want_h, want_c = O.mesh_batch_hash(vox0, nbr_np, mode)
            ok = np.array_equal(host.meta[:, 1], want_c) and np.array_equal(got_h, want_h)
            checked = n_step
        parity = {"ok": all_ok(bool(ok)), "chunks_checked_per_rank": int(checked),
                  "what": "per-chunk quad counts and SPEC §4 key hashes of every rank's CUDA result vs the in-repo CPU oracle"}

    # ---------------- warm-up (>= W steps and >= 0.3 s so clocks have ramped)
    sampler = ClockSampler(local)
    sampler.start()
    t_w = time.time()
    k = 0
    while k < args.warmup or time.time() - t_w < args.soak_seconds:
        launch(k, outs[k % 2])
        k += 1
        if k % 64 == 0:
            torch.cuda.synchronize()
    # the gather's torch kernels (strided slice -> int64, clone / all_gather) are loaded lazily by the driver on
    # their first launch (tens of ms): run them once here so that cost is not inside the timed region
    n_total = n_step * world if not fused else coords_np.shape[0]
    _ = gather_counts(out.meta[:, 1].to(torch.int64), n_total, rank, world)
    barrier()

    # ---------------- timed region: exactly K steps, back to back on the launching stream, device-timed
    launches0 = m.launches
    e_start, e_kend, e_end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.time()
    e_start.record()
    for s in range(args.steps):
        launch(s, outs[s % 2])
    e_kend.record()
    out = outs[(args.steps - 1) % 2]
    # the one collective of the path: final host gather of mesh sizes (per-chunk quad counts)
    final_counts = gather_counts(out.meta[:, 1].to(torch.int64), n_total, rank, world)
    e_end.record()
    barrier()
    t1 = time.time()
    launches = m.launches - launches0
    elapsed_ms = max_over_ranks(e_start.elapsed_time(e_end))
    ms_per_step = elapsed_ms / args.steps
    clocks = sampler.stop(t0, t1, "timed region")
    units_per_step = n_total
    value = units_per_step / (ms_per_step * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / average launch duration over the timed
    # region (events around the K launches, gather excluded).  With overlap on, consecutive launches share the GPU
    # for the length of a tail, so the average duration is the launch-to-launch period; the duration of a launch
    # that has the GPU to itself is measured separately below (overlap off, every launch between its own events).
    mean_ms = e_start.elapsed_time(e_kend) / args.steps
    mean_bytes = statistics.mean(alg_bytes[s % len(batches)] for s in range(args.steps))
    m.set_overlap(False)
    iso_n = min(args.steps, 140)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iso_n)]
    for s in range(iso_n):
        ev[s][0].record()
        launch(s, outs[s % 2])
        ev[s][1].record()
    torch.cuda.synchronize()
    m.set_overlap(overlap)
    iso_ms = [a.elapsed_time(b) for a, b in ev]
    iso_bytes = statistics.mean(alg_bytes[s % len(batches)] for s in range(iso_n))
    achieved = mean_bytes / (mean_ms * 1e-3) / 1e9
    iso_achieved = iso_bytes / (statistics.mean(iso_ms) * 1e-3) / 1e9
    kernel_name = ("mesh::mesh_kernel_g" if mode == GREEDY else "mesh::mesh_kernel_c") if not fused else "terrain_mesh_kernel<true>"
    traffic, traffic_src = ncu_traffic(kernel_name)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_isolated": iso_achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kernel_name,
                "kernel_ms_mean": mean_ms, "algorithmic_bytes_per_launch": mean_bytes, "launches_timed": args.steps,
                "overlap": overlap,
                "isolated": {"kernel_ms_mean": statistics.mean(iso_ms), "kernel_ms_min": min(iso_ms), "launches": iso_n,
                             "achieved": iso_achieved, "frac": iso_achieved / peak,
                             "note": "overlap off, each launch between its own pair of events"},
                "note": "frac: launch-to-launch period of K back-to-back launches (consecutive launches of a context overlap: "
                        "vxp_ctx_set_overlap); frac_isolated: one launch alone on the GPU; no memset, no other kernel"}
    if fused:
        # the fused path touches HBM only for its output; SURVEY §8d asks for the labelled figure against the bytes
        # the unfused path would move (terrain written, voxels read back, mesh written) so fusion is not mistaken for bandwidth
        unfused_bytes = n_step * (2 * CHUNK_BYTES + 8) + 56 * int(counts[0].sum())
        roofline["B_terrain+B_mesh"] = {"bytes_per_launch": unfused_bytes, "equivalent_gbs": unfused_bytes / (mean_ms * 1e-3) / 1e9,
                                         "equivalent_frac": unfused_bytes / (mean_ms * 1e-3) / 1e9 / peak,
                                         "note": "what the unfused path (fill 64 KiB, read 64 KiB, write the mesh) would have moved "
                                                 "in the same time; not a bandwidth claim: the fused kernel is ALU/latency-bound"}

    # ---------------- e2e: host buffers in, host mesh out, through vxp_mesh_chunks_host / vxp_terrain_mesh_host
    e2e = None
    if args.e2e_steps > 0:
        import ctypes as C
        lib = m._lib
        if fused:
            h_coords = np.ascontiguousarray(coords_np[lo:hi])
            call = lambda s: m.terrain_mesh_host(h_coords, 42, mode, copy=False)
            h2d = h_coords.nbytes
        else:
            pinned = []
            for b in batches[:2]:                       # two pinned 256 MiB host batches alternate
                p = C.c_void_p()
                assert lib.vxp_host_alloc(n_step * CHUNK_BYTES, C.byref(p)) == 0
                torch.cuda.synchronize()
                hb = np.frombuffer((C.c_char * (n_step * CHUNK_BYTES)).from_address(p.value), dtype=np.uint16)
                hb[:] = b.cpu().numpy().view(np.uint16).reshape(-1)
                pinned.append(p)
            h_nbr = np.ascontiguousarray(nbr_np)
            call = lambda s: m.mesh_host_ptr(pinned[s % 2].value, h_nbr.ctypes.data, n_step, mode)
            h2d = n_step * CHUNK_BYTES + h_nbr.nbytes
        m.use_stream(None)
        for s in range(2):
            res = call(s)
        barrier()
        tw0 = time.perf_counter()
        d2h = 0
        for s in range(args.e2e_steps):
            res = call(s)
            d2h += 8 + 8 * n_step + 56 * res.total_quads
        barrier()
        dt = max_over_ranks(time.perf_counter() - tw0)
        e2e = {"value": units_per_step * args.e2e_steps / dt, "unit": "chunks/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h // args.e2e_steps),
               "steps": args.e2e_steps, "api": "vxp_terrain_mesh_host" if fused else "vxp_mesh_chunks_host",
               "host_memory": "pinned (vxp_host_alloc)"}
        if not fused:
            # the host link itself: the same pinned 256 MiB batch copied up with nothing else going on (per GPU, all
            # ranks at the same time) - the ceiling of any call that has to move the raw voxels
            dst = torch.empty(n_step * CHUNK_BYTES, dtype=torch.uint8, device=dev)
            src = torch.from_numpy(np.frombuffer((C.c_char * (n_step * CHUNK_BYTES)).from_address(pinned[0].value), dtype=np.uint8))
            dst.copy_(src, non_blocking=True)
            barrier()
            tc0 = time.perf_counter()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            barrier()
            dtc = max_over_ranks(time.perf_counter() - tc0)
            link = 4 * n_step * CHUNK_BYTES / dtc / 1e9
            e2e["h2d_link_gbs_per_gpu_all_ranks_busy"] = link
            e2e["upload_bound_chunks_per_s"] = world * link * 1e9 / CHUNK_BYTES
            e2e["frac_of_upload_bound"] = e2e["value"] / e2e["upload_bound_chunks_per_s"]
            del dst
            if not args.no_also:
                # the same call from PAGEABLE caller memory (what a Vec<u16> is): the driver stages the copy itself
                pg = [np.array(np.frombuffer((C.c_char * (n_step * CHUNK_BYTES)).from_address(p.value), dtype=np.uint16)) for p in pinned[:1]]
                for _ in range(2):
                    m.mesh_host_ptr(pg[0].ctypes.data, h_nbr.ctypes.data, n_step, mode)
                barrier()
                tp0 = time.perf_counter()
                for _ in range(4):
                    m.mesh_host_ptr(pg[0].ctypes.data, h_nbr.ctypes.data, n_step, mode)
                barrier()
                e2e["pageable_chunks_per_s"] = units_per_step * 4 / max_over_ranks(time.perf_counter() - tp0)
                del pg

                # lighter ways through the same host API, beside the number above and never instead of it; all ranks
                # at the same time, wall clock, max over ranks
                def all_ranks(fn, reps=6):
                    for _ in range(2):
                        fn()
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        r = fn()
                    barrier()
                    return units_per_step * reps / max_over_ranks(time.perf_counter() - t0), r

                light = {}
                # (1) no index buffer in the result (SPEC §5: indices are 4q + {0,1,2,2,3,0}): 24 of 56 bytes per quad stay home
                m.set_host_indices(False)
                v, r = all_ranks(lambda: m.mesh_host_ptr(pinned[0].value, h_nbr.ctypes.data, n_step, mode))
                m.set_host_indices(True)
                light["vertices_only"] = {"chunks_per_s": v, "d2h_bytes": int(8 + 8 * n_step + 32 * r.total_quads)}
                # (2) packed upload (SPEC §10): the batch packed once on the device, outside the timed region, as a server
                # that stores chunks packed would have it; packed bytes up, chunks expanded on the device
                m.use_current_torch_stream()
                pk, po, pt = m.pack(batches[0])
                torch.cuda.synchronize()
                pbytes = int(pt.item())
                pp, op = C.c_void_p(), C.c_void_p()
                assert lib.vxp_host_alloc(pbytes, C.byref(pp)) == 0 and lib.vxp_host_alloc(8 * n_step, C.byref(op)) == 0
                np.frombuffer((C.c_char * pbytes).from_address(pp.value), dtype=np.uint8)[:] = pk[:pbytes].cpu().numpy()
                np.frombuffer((C.c_char * (8 * n_step)).from_address(op.value), dtype=np.int64)[:] = po.cpu().numpy()
                m.use_stream(None)
                v, r = all_ranks(lambda: m.mesh_packed_host_ptr(pp.value, pbytes, op.value, h_nbr.ctypes.data, n_step, mode))
                light["packed_upload"] = {"chunks_per_s": v, "h2d_bytes": int(pbytes + 8 * n_step + h_nbr.nbytes),
                                          "d2h_bytes": int(8 + 8 * n_step + 56 * r.total_quads), "api": "vxp_mesh_packed_host"}
                m.set_host_indices(False)
                v, r = all_ranks(lambda: m.mesh_packed_host_ptr(pp.value, pbytes, op.value, h_nbr.ctypes.data, n_step, mode))
                m.set_host_indices(True)
                light["packed_upload_vertices_only"] = {"chunks_per_s": v, "d2h_bytes": int(8 + 8 * n_step + 32 * r.total_quads)}
                lib.vxp_host_free(pp)
                lib.vxp_host_free(op)
                del pk, po
                # (3) from chunk coordinates: terrain generated on the device, 12 bytes per chunk go up
                hc = np.ascontiguousarray(coords_np)
                v, r = all_ranks(lambda: m.terrain_mesh_host(hc, 42 + rank, mode, copy=False))
                light["from_coords"] = {"chunks_per_s": v, "h2d_bytes": int(hc.nbytes),
                                        "d2h_bytes": int(8 + 8 * n_step + 56 * r.total_quads), "api": "vxp_terrain_mesh_host"}
                e2e["lighter_paths"] = light
        m.use_current_torch_stream()

    # ---------------- secondary measurements (not the headline; same hygiene, fewer steps)
    also = {}

    def timed(fn, steps=60):
        for s in range(5):
            fn(s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(steps):
            fn(s)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    if not args.no_also and not fused:
        # the other mode on the same batches, overlapped and one launch at a time
        other = GREEDY if mode == CULLED else CULLED
        oprobe = m.alloc_mesh(n_step, 0)
        m.set_overlap(False)
        ocounts = []
        for k in range(len(batches)):
            m.mesh(batches[k], dnbr, other, out=oprobe)
            torch.cuda.synchronize()
            ocounts.append(oprobe.meta[:, 1].cpu().numpy().astype(np.int64))
        oouts = [m.alloc_mesh(n_step, max(int(c.sum()) for c in ocounts) + 1) for _ in range(2)]
        ms_iso = max_over_ranks(timed(lambda s: m.mesh(batches[s % len(batches)], dnbr, other, out=oouts[s % 2])))
        m.set_overlap(overlap)
        ms = max_over_ranks(timed(lambda s: m.mesh(batches[s % len(batches)], dnbr, other, out=oouts[s % 2]), steps=200))
        m.set_overlap(False)
        ob = statistics.mean(algorithmic_bytes(c) for c in ocounts)
        also["greedy" if other == GREEDY else "culled"] = {
            "chunks_per_s": n_step * world / (ms * 1e-3), "ms_per_step": ms, "hbm_gbs_per_gpu": ob / (ms * 1e-3) / 1e9,
            "frac": ob / (ms * 1e-3) / 1e9 / peak, "overlap": overlap, "ms_per_step_isolated": ms_iso,
            "frac_isolated": ob / (ms_iso * 1e-3) / 1e9 / peak,
            "quads_per_batch": float(np.mean([c.sum() for c in ocounts]))}
        del oouts, oprobe

    if not args.no_also and not args.workload.startswith("checker"):
        # config 5 (the north star's multi-GPU configuration): the 65^3 world, fused terrain + greedy, strong-scaled
        # over the ranks by coordinate ranges; one step = the whole world
        wc = grid_coords_np(WORLD_LO, WORLD_HI)
        wlo, whi = partition(wc.shape[0], rank, world)
        wmine = torch.from_numpy(np.ascontiguousarray(wc[wlo:whi])).to(dev)
        wprobe = m.alloc_mesh(whi - wlo, 0)
        m.terrain_mesh(wmine, 42, GREEDY, out=wprobe)
        torch.cuda.synchronize()
        wq = int(wprobe.total.item())
        wout = m.alloc_mesh(whi - wlo, wq + 1)
        barrier()
        wms = max_over_ranks(timed(lambda s: m.terrain_mesh(wmine, 42, GREEDY, out=wout), steps=30))
        wtot = torch.tensor([wq], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(wtot)
        unf = wc.shape[0] * (2 * CHUNK_BYTES + 8) + 56 * int(wtot.item())
        also["world_fused"] = {"chunks_per_s": wc.shape[0] / (wms * 1e-3), "ms_per_world": wms, "scaling": "strong",
                               "chunks": int(wc.shape[0]), "quads": int(wtot.item()),
                               "B_terrain+B_mesh_equivalent_gbs_per_gpu": unf / world / (wms * 1e-3) / 1e9,
                               "note": "BASELINE config 5; also the headline of --workload world_fused"}
        del wout, wprobe, wmine

    if not args.no_also and not fused and world == 1:
        # the same workload with two batches in flight (two contexts, two streams): the tail of one launch and the
        # ramp of the next overlap, which a single in-order stream cannot do
        m2 = Mesher(local)
        s2 = torch.cuda.Stream(device=dev)
        m2.use_stream(s2)
        out2 = m.alloc_mesh(n_step, max(totals) + 1)
        pair = [(m, out, stream), (m2, out2, s2)]

        def two(s):
            mm, oo, _ = pair[s % 2]
            mm.mesh(batches[s % len(batches)], dnbr, mode, out=oo)

        for s in range(6):
            two(s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        s2.wait_event(a)
        for s in range(120):
            two(s)
        j = torch.cuda.Event()
        j.record(s2)
        stream.wait_event(j)
        b.record(stream)
        torch.cuda.synchronize()
        ms2 = a.elapsed_time(b) / 120
        also["two_batches_in_flight"] = {"chunks_per_s": n_step / (ms2 * 1e-3), "ms_per_step": ms2,
                                         "hbm_gbs": mean_bytes / (ms2 * 1e-3) / 1e9, "frac": mean_bytes / (ms2 * 1e-3) / 1e9 / peak,
                                         "note": "two contexts on two streams alternate over the same rotating batches"}
        m2.close()
        if not args.workload.startswith("checker"):
            scratch = torch.empty_like(batches[0])
            ms = timed(lambda s: m.terrain_fill(dcoords, 100 + s, out=scratch))
            also["terrain_fill"] = {"chunks_per_s": n_step / (ms * 1e-3), "ms_per_step": ms,
                                    "hbm_gbs": n_step * CHUNK_BYTES / (ms * 1e-3) / 1e9,
                                    "frac": n_step * CHUNK_BYTES / (ms * 1e-3) / 1e9 / peak}
            # a world-sized batch (32^3 chunks, 2 GiB of voxels): the chunks of a column share its heights
            big_c = torch.from_numpy(grid_coords_np((0, 0, 0), (32, 32, 32))).to(dev)
            big_v = torch.empty((32768, 32768), dtype=torch.int16, device=dev)
            ms = timed(lambda s: m.terrain_fill(big_c, 100 + s, out=big_v), steps=10)
            also["terrain_fill_32768"] = {"chunks_per_s": 32768 / (ms * 1e-3), "ms_per_step": ms,
                                          "hbm_gbs": 32768 * CHUNK_BYTES / (ms * 1e-3) / 1e9,
                                          "frac": 32768 * CHUNK_BYTES / (ms * 1e-3) / 1e9 / peak}
            del big_v
            fprobe = m.alloc_mesh(n_step, 0)
            m.terrain_mesh(dcoords, 42, GREEDY, out=fprobe)
            torch.cuda.synchronize()
            fout = m.alloc_mesh(n_step, int(fprobe.total.item()) + 1)
            ms = timed(lambda s: m.terrain_mesh(dcoords, 42, GREEDY, out=fout))
            also["fused_terrain_greedy"] = {"chunks_per_s": n_step / (ms * 1e-3), "ms_per_step": ms,
                                            "quads": int(fprobe.total.item())}
            # the whole north-star path through the host API: chunk coordinates in, host meshes out (terrain is
            # generated on the device, so only 12 bytes per chunk go up; the mesh download bounds the call)
            hc = np.ascontiguousarray(coords_np)
            m.use_stream(None)
            for _ in range(2):
                r = m.terrain_mesh_host(hc, 42, mode, copy=False)
            tw = time.perf_counter()
            reps = 8
            for _ in range(reps):
                r = m.terrain_mesh_host(hc, 42, mode, copy=False)
            dtw = (time.perf_counter() - tw) / reps
            also["e2e_from_coords"] = {"chunks_per_s": n_step / dtw, "ms_per_call": dtw * 1e3, "api": "vxp_terrain_mesh_host",
                                       "h2d_bytes": int(hc.nbytes), "d2h_bytes": int(8 + 8 * n_step + 56 * r.total_quads)}
            # the same batch through the packed upload path (SPEC §10): pack once on the device (outside the
            # timed region, as a server that stores chunks packed would have them), then time host packed bytes in,
            # host meshes out; plus the pack / unpack kernels on their own (bytes of raw voxels per second)
            import ctypes as C2
            m.use_current_torch_stream()
            pk, po, pt = m.pack(batches[0])
            torch.cuda.synchronize()
            pbytes = int(pt.item())
            pp, op = C2.c_void_p(), C2.c_void_p()
            assert m._lib.vxp_host_alloc(pbytes, C2.byref(pp)) == 0 and m._lib.vxp_host_alloc(8 * n_step, C2.byref(op)) == 0
            np.frombuffer((C2.c_char * pbytes).from_address(pp.value), dtype=np.uint8)[:] = pk[:pbytes].cpu().numpy()
            np.frombuffer((C2.c_char * (8 * n_step)).from_address(op.value), dtype=np.int64)[:] = po.cpu().numpy()
            hn2 = np.ascontiguousarray(nbr_np)
            m.use_stream(None)
            for _ in range(3):
                r = m.mesh_packed_host_ptr(pp.value, pbytes, op.value, hn2.ctypes.data, n_step, mode)
            tw = time.perf_counter()
            for _ in range(8):
                r = m.mesh_packed_host_ptr(pp.value, pbytes, op.value, hn2.ctypes.data, n_step, mode)
            dtw = (time.perf_counter() - tw) / 8
            m.use_current_torch_stream()
            ms_pack = timed(lambda s: m.pack(batches[s % len(batches)]), steps=20)
            unp = torch.empty_like(batches[0])
            ms_unpack = timed(lambda s: m.unpack(pk, po, out=unp), steps=40)
            m.synchronize()
            also["e2e_packed"] = {"chunks_per_s": n_step / dtw, "ms_per_call": dtw * 1e3, "api": "vxp_mesh_packed_host",
                                  "h2d_bytes": int(pbytes + 8 * n_step + hn2.nbytes), "raw_bytes": n_step * CHUNK_BYTES,
                                  "d2h_bytes": int(8 + 8 * n_step + 56 * r.total_quads),
                                  "pack_kernel_ms": ms_pack, "unpack_kernel_ms": ms_unpack,
                                  "pack_frac": (n_step * CHUNK_BYTES + pbytes) / (ms_pack * 1e-3) / 1e9 / peak,
                                  "unpack_frac": (n_step * CHUNK_BYTES + pbytes) / (ms_unpack * 1e-3) / 1e9 / peak}
            m._lib.vxp_host_free(pp)
            m._lib.vxp_host_free(op)
            # level of detail: the 16^3 block of chunks -> 8^3 half-resolution chunks (reads 8 bytes per byte written)
            k8 = np.array([[(2 * X + (c & 1)) + 16 * ((2 * Y + ((c >> 1) & 1)) + 16 * (2 * Z + (c >> 2))) for c in range(8)]
                           for Z in range(8) for Y in range(8) for X in range(8)], dtype=np.int32)
            dk8 = torch.from_numpy(k8).to(dev)
            lod_out = torch.empty((512, 32768), dtype=torch.int16, device=dev)
            ms_lod = timed(lambda s: m.downsample(batches[s % len(batches)], dk8, out=lod_out), steps=40)
            also["lod_downsample"] = {"ms_per_step": ms_lod, "chunks_in": n_step, "chunks_out": 512,
                                      "hbm_gbs": (n_step + 512) * CHUNK_BYTES / (ms_lod * 1e-3) / 1e9,
                                      "frac": (n_step + 512) * CHUNK_BYTES / (ms_lod * 1e-3) / 1e9 / peak}
            # latency of one voxel edit -> re-meshed chunks, device level (no host round trip inside;
            # timed with events over back-to-back calls) and host level (edits up, new meshes down, wall clock)
            from voxelphile_b200 import EDIT_DTYPE
            surf = int(np.argmax(counts[0]))                                 # a surface chunk of batch 0
            ed = np.array([(surf, 0, 16, 31, 0)], dtype=EDIT_DTYPE)          # on two faces: three chunks re-meshed
            ded = torch.from_numpy(ed.view(np.int32).reshape(-1, 3)).to(dev)
            m.use_current_torch_stream()
            work = batches[0].clone()
            for _ in range(5):
                dl, dcn, dm = m.remesh_edits(work, dnbr, ded, mode)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(200):
                dl, dcn, dm = m.remesh_edits(work, dnbr, ded, mode)
            b.record()
            torch.cuda.synchronize()
            dev_us = a.elapsed_time(b) / 200 * 1e3
            m.use_stream(None)
            hv0 = batches[0].cpu().numpy().view(np.uint16)
            m.mesh_host(hv0, nbr_np, mode, copy=False)
            for _ in range(3):
                m.remesh_edits_host(ed, mode, copy=False)
            tw = time.perf_counter()
            for _ in range(200):
                dirty, hm = m.remesh_edits_host(ed, mode, copy=False)
            host_us = (time.perf_counter() - tw) / 200 * 1e6
            m.use_current_torch_stream()
            also["remesh_one_edit"] = {"device_us": dev_us, "host_us": host_us, "chunks_remeshed": int(len(dirty)),
                                       "quads": int(hm.total_quads),
                                       "note": "vxp_remesh_edits / vxp_remesh_edits_host, one edit on a chunk edge"}

    # ---------------- CPU baseline: the oracle on this box's host cores (rank 0, N=1 only), same batch and tables
    cpu = None
    if rank == 0 and world == 1 and not fused and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        threads = O.max_threads()
        hv = batches[0].cpu().numpy().view(np.uint16)
        hn = nbr_np
        cap = int(O.mesh_batch_hash(hv, hn, mode, threads)[1].sum())
        best = None
        t_begin = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t_begin < 10 and reps < 20):
            tc = time.perf_counter()
            O.mesh_batch(hv, hn, mode, threads, capacity=cap)
            dtc = time.perf_counter() - tc
            best = dtc if best is None else min(best, dtc)
            reps += 1
        if "remesh_one_edit" in also:
            # the same three chunks through the oracle, one thread (what a CPU client would do per edit)
            k3 = [surf] + [int(nbr_np[surf][f]) for f in (0, 5) if nbr_np[surf][f] >= 0]
            tc = time.perf_counter()
            for _ in range(20):
                for c in k3:
                    O.mesh_chunk_keys(hv[c], [None if hn[c, f] < 0 else hv[hn[c, f]] for f in range(6)], mode)
            also["remesh_one_edit"]["cpu_oracle_us"] = (time.perf_counter() - tc) / 20 * 1e6
        # SURVEY §8d also asks for the single-thread figure: the first 512 chunks (z-layers 0 and 1: every height band)
        n1 = 512
        hn1 = hn[:n1].copy()
        hn1[hn1 >= n1] = -1
        cap1 = int(O.mesh_batch_hash(hv[:n1], hn1, mode, 1)[1].sum())
        t1best = None
        for _ in range(3):
            tc = time.perf_counter()
            O.mesh_batch(hv[:n1], hn1, mode, 1, capacity=cap1)
            dtc = time.perf_counter() - tc
            t1best = dtc if t1best is None else min(t1best, dtc)
        cpu = {"value": n_step / best, "unit": "chunks/s", "cores": threads, "kind": "port",
               "value_1_thread": n1 / t1best,
               "sample": f"the whole 4096-chunk batch 0 (seed 42) with its neighbour table, best of {reps} passes, "
                         f"in-repo C oracle (scalar, pthreads); the reference has no mesher to time"}

    if rank == 0:
        line = {
            "metric": "chunks meshed/sec", "value": value, "unit": "chunks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": args.workload, "chunks_per_step_per_gpu": n_step, "chunk": "32^3 u16",
                       "mode": "greedy" if mode == GREEDY else "culled",
                       "batches": f"{len(batches)} rotating batches of {n_step} chunks "
                                  f"({len(batches) * n_step * CHUNK_BYTES >> 20} MiB > 126 MB L2)" if not fused else
                                  "fused terrain+mesh, no stored voxels",
                       "quads_per_step_per_gpu": float(np.mean([c.sum() for c in counts])),
                       "voxels_per_s": value * 32768, "spec": "SPEC.md v1 (builder-defined; parity unpinned)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "gathered_quads": int(final_counts.sum().item()),
            "parity_check": ("ok" if parity["ok"] else "FAILED") if parity else None, "parity": parity, "also": also,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    m.close()


if __name__ == "__main__":
    main()
