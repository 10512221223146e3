"""Host-to-device bandwidth of one 142 MB fp32 batch from pinned memory: one copy, the same split over several streams, and
write-combined pinned memory.  (e2e of bench.py is bound by this copy: 27-29 GB/s on the test boxes.)"""
import ctypes, sys, time
import torch
n = 32 * 3 * 608 * 608
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.uniform_()
dev = torch.empty(n, dtype=torch.float32, device="cuda")
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = timed(lambda: dev.copy_(host, non_blocking=True))
print(f"1 stream : {n * 4 / t / 1e9:6.1f} GB/s")
for k in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    hs, ds = host.chunk(k), dev.chunk(k)
    def multi():
        for s, h, d in zip(streams, hs, ds):
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
    t = timed(multi)
    print(f"{k} streams: {n * 4 / t / 1e9:6.1f} GB/s")
rt = torch.cuda.cudart()
libc = ctypes.CDLL("libcudart.so", mode=ctypes.RTLD_GLOBAL) if False else None
try:
    import numpy as np
    lib = ctypes.CDLL(next(p for p in [l.split()[-1] for l in open("/proc/self/maps") if "libcudart" in l]))
    ptr = ctypes.c_void_p()
    rc = lib.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n * 4), ctypes.c_uint(4))    # cudaHostAllocWriteCombined
    assert rc == 0, rc
    arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n,))
    arr[:] = 0.5
    wc = torch.from_numpy(arr)
    lib.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    st = torch.cuda.current_stream().cuda_stream
    t = timed(lambda: lib.cudaMemcpyAsync(dev.data_ptr(), ptr, n * 4, 1, st))
    print(f"write-combined pinned, 1 stream: {n * 4 / t / 1e9:6.1f} GB/s")
except Exception as e:
    print("write-combined probe failed:", type(e).__name__, e)
# device-to-host for completeness
t = timed(lambda: host.copy_(dev, non_blocking=True))
print(f"D2H 1 stream: {n * 4 / t / 1e9:6.1f} GB/s")
