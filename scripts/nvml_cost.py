import time, pynvml, torch
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.zeros(1<<20, device="cuda"); torch.cuda.synchronize()
for name, fn in (("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ("maxclock", lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))):
    ts = []
    for _ in range(12):
        t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
    print(name, [round(v, 2) for v in ts])
