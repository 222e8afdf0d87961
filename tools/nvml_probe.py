import time, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
for f in ('nvmlDeviceGetClockInfo', 'nvmlDeviceGetCurrentClocksEventReasons'):
    fn = getattr(pynvml, f)
    t0 = time.perf_counter()
    for _ in range(20):
        fn(h, pynvml.NVML_CLOCK_SM) if 'ClockInfo' in f else fn(h)
    print(f, (time.perf_counter() - t0) / 20 * 1e3, 'ms/call')
