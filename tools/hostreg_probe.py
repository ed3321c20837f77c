"""Host-side probe for the e2e leg: what does DMA straight from the page cache cost?  mmap a tmpfs file, cudaHostRegister windows of it
(read-only), time the registration and the H2D copy from the registered window, against pread -> pinned -> H2D (the loader's path).
Output: one JSON object on stdout.  Needs one GPU; torch-free (libcudart through ctypes)."""
import json
import mmap
import os
import sys
import threading
import time

import ctypes

N_GB = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
WIN_MB = [16, 64, 256, 1024]
path = "/dev/shm/kk_hostreg_probe.bin"
size = int(N_GB * (1 << 30))
out = {"file_GB": size / 1e9}
t0 = time.time()
with open(path, "wb") as f:
    blk = os.urandom(1 << 24)
    for _ in range(size >> 24):
        f.write(blk)
out["write_s"] = time.time() - t0
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
assert rt.cudaSetDevice(0) == 0
_dev = ctypes.c_void_p()
assert rt.cudaMalloc(ctypes.byref(_dev), size) == 0
dev_ptr = _dev.value


def sync():
    assert rt.cudaDeviceSynchronize() == 0


fd = os.open(path, os.O_RDONLY)
mm = mmap.mmap(fd, size, mmap.MAP_SHARED, mmap.PROT_READ)
import numpy as np  # noqa: E402

arr = np.frombuffer(mm, dtype=np.uint8)  # address of the mapping without copying
addr = arr.ctypes.data
cudaHostRegisterReadOnly, cudaHostRegisterDefault = 0x08, 0x0


def reg(a, n, flags):
    return int(rt.cudaHostRegister(a, n, flags))


def unreg(a):
    return int(rt.cudaHostUnregister(a))


def h2d(dst_off, src_addr, n, stream):
    # cudaMemcpyAsync through torch's cudart binding
    return int(rt.cudaMemcpyAsync(dev_ptr + dst_off, src_addr, n, 1, stream))


res = {}
for flags, fname in ((cudaHostRegisterReadOnly, "readonly"), (cudaHostRegisterDefault, "default")):
    for wmb in WIN_MB:
        w = wmb << 20
        nwin = min(size // w, max(1, (2 << 30) // w))
        t_reg = t_copy = 0.0
        ok = True
        s = None
        for i in range(nwin):
            a = addr + i * w
            t = time.perf_counter()
            rc = reg(a, w, flags)
            t_reg += time.perf_counter() - t
            if rc != 0:
                ok = False
                res[f"{fname}_{wmb}MB"] = {"error": f"cudaHostRegister rc={rc}"}
                break
            t = time.perf_counter()
            h2d(i * w, a, w, s)
            sync()
            t_copy += time.perf_counter() - t
            unreg(a)
        if ok:
            tot = nwin * w
            res[f"{fname}_{wmb}MB"] = {"register_GBps": tot / t_reg / 1e9, "h2d_GBps": tot / t_copy / 1e9, "serial_GBps": tot / (t_reg + t_copy) / 1e9, "windows": nwin}
out["register_then_copy"] = res

# threads registering disjoint windows concurrently (does the mm lock serialise them?)
for nthr in (4, 16):
    w = 64 << 20
    per = max(1, (size // w) // nthr)
    errs = []

    def work(k):
        for i in range(per):
            a = addr + (k * per + i) * w
            if reg(a, w, cudaHostRegisterReadOnly) != 0:
                errs.append(1)
                return
    t = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(nthr)]
    [x.start() for x in th]
    [x.join() for x in th]
    dt = time.perf_counter() - t
    out[f"parallel_register_{nthr}thr_GBps"] = None if errs else nthr * per * w / dt / 1e9
    if not errs:  # whole range registered: one big copy
        tot = nthr * per * w
        sync()
        t = time.perf_counter()
        h2d(0, addr, tot, None)
        sync()
        out[f"h2d_from_registered_pagecache_{nthr}thr_GBps"] = tot / (time.perf_counter() - t) / 1e9
    for k in range(nthr):
        for i in range(per):
            unreg(addr + (k * per + i) * w)

# the loader's path for comparison: pread into a pinned buffer, then H2D
_pin = ctypes.c_void_p()
assert rt.cudaHostAlloc(ctypes.byref(_pin), 64 << 20, 0) == 0
pv = np.ctypeslib.as_array(ctypes.cast(_pin.value, ctypes.POINTER(ctypes.c_uint8)), shape=(64 << 20,))
t = time.perf_counter()
tot = 0
for i in range(min(size // (64 << 20), 32)):
    os.preadv(fd, [memoryview(pv)], i * (64 << 20))
    tot += 64 << 20
out["pread_single_thread_GBps"] = tot / (time.perf_counter() - t) / 1e9
os.remove(path)
print(json.dumps(out))
