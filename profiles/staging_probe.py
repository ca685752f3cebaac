"""Host <-> device staging probe (round 2): where do the seconds of the e2e path go on this host?

    python profiles/staging_probe.py [GB=8] > profiles/staging_probe_r2.json

Measures, for pool sizes 4..96 threads: host fingerprint rate (read-only pass), staged H2D of a pageable int64 array
with narrowing and of a float32 array, staged D2H into a fresh / an already touched destination; plus the raw DMA
rates pinned<->device and the cost of freeing a large host array.
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from muon_b200 import _device  # noqa: E402

GB = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(GB * 2**30 / 8)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
idx64 = rng.integers(0, 200_000, n, dtype=np.int64)            # GB gigabytes of int64 indices
val32 = rng.random(n, dtype=np.float32)                        # GB/2 of float32
out = {"gb_int64": idx64.nbytes / 2**30, "gb_f32": val32.nbytes / 2**30, "threads": {}}

# raw DMA rates from / to pinned memory
pin = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for name, a, b in (("dma_h2d_GBps", pin, d), ("dma_d2h_GBps", d, pin)):
    b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    out[name] = 8 * (1 << 30) / (time.perf_counter() - t0) / 1e9
del pin, d

d_idx = torch.empty(n, dtype=torch.int32, device=dev)
d_val = torch.empty(n, dtype=torch.float32, device=dev)
touched = np.zeros(n, dtype=np.float32)
for T in (4, 8, 16, 32, 48, 64, 96):
    st = _device.Stager(threads=T)
    r = {}
    t0 = time.perf_counter(); st.fingerprint(idx64); r["fingerprint_int64_GBps"] = idx64.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); st.fingerprint(val32); r["fingerprint_f32_GBps"] = val32.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); st.h2d(idx64, d_idx, narrow=True); torch.cuda.synchronize()
    r["h2d_narrow_out_GBps"] = d_idx.numel() * 4 / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); st.h2d(val32, d_val); torch.cuda.synchronize()
    r["h2d_f32_GBps"] = val32.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); st.d2h(d_val, touched); r["d2h_touched_GBps"] = val32.nbytes / (time.perf_counter() - t0) / 1e9
    fresh = np.empty(n, dtype=np.float32)
    t0 = time.perf_counter(); st.d2h(d_val, fresh); r["d2h_fresh_GBps"] = val32.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); del fresh; r["free_fresh_s_per_GB"] = (time.perf_counter() - t0) / (val32.nbytes / 2**30)
    out["threads"][T] = {k: round(v, 3) for k, v in r.items()}
    del st
dfp = _device.device_fingerprint(d_idx)
out["device_fingerprint_equals_host"] = bool(dfp == _device.host_pool().fingerprint(idx64))
t0 = time.perf_counter(); _device.device_fingerprint(d_idx); torch.cuda.synchronize()
out["device_fingerprint_GBps"] = d_idx.numel() * 4 / (time.perf_counter() - t0) / 1e9
print(json.dumps(out, indent=1))
