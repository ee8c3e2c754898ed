"""How long do the NVML queries of bench.py's ClockSampler take? (they contend with kernel launches for driver locks)"""
import time, pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
for name, fn in [("clock", lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                 ("reasons", lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ("util", lambda: nv.nvmlDeviceGetUtilizationRates(h).gpu)]:
    ts = []
    for _ in range(20):
        t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
    print(f"{name:8s} min {min(ts):.3f} ms  median {sorted(ts)[10]:.3f} ms  max {max(ts):.3f} ms")
