"""Host->device staging probe: what limits the upload of scipy's int64 index array (48 GB at the headline config)?
python profiles/h2d_probe.py [elements]   -- prints GB/s of SOURCE bytes for each variant."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muon_b200 import _device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
src = np.arange(n, dtype=np.int64) % 200_000            # pageable, touched
dev = torch.device("cuda")
gb = src.nbytes / 1e9


def T(label, fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
        del r
    print(f"{label:58s} {best:7.3f} s  {gb / best:6.1f} GB/s of source")


# PCIe peak from pinned memory
pin = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(8):
    dst.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
print(f"pinned 1 GiB x8 H2D: {8 * (1 << 30) / 1e9 / (time.perf_counter() - t):.1f} GB/s")

for threads in (16, 32, 64):
    _device._COPY_THREADS = threads
    _device._POOL = ThreadPoolExecutor(threads)
    T(f"to_device(int64 -> int32 on device), {threads} copy threads", lambda: _device.to_device(src, dev, np.int32))


def host_cast(threads):
    """narrow to int32 while copying into the pinned stage (halves the PCIe bytes)."""
    pool = ThreadPoolExecutor(threads)
    stage = [torch.empty(64 << 20, dtype=torch.int32).pin_memory() for _ in range(2)]
    out = torch.empty(n, dtype=torch.int32, device=dev)
    step = stage[0].numel()
    ev = [None, None]
    for i, off in enumerate(range(0, n, step)):
        m = min(step, n - off)
        if ev[i & 1] is not None:
            ev[i & 1].synchronize()
        s = stage[i & 1][:m].numpy()
        b = [m * j // threads for j in range(threads + 1)]
        list(pool.map(lambda j: np.copyto(s[b[j]:b[j + 1]], src[off + b[j]:off + b[j + 1]], casting="unsafe"), range(threads)))
        out[off:off + m].copy_(stage[i & 1][:m], non_blocking=True)
        e = torch.cuda.Event()
        e.record()
        ev[i & 1] = e
    return out


for threads in (16, 32, 64):
    T(f"int64 -> int32 on the host while staging, {threads} threads", lambda: host_cast(threads))
chk = host_cast(32)
assert torch.equal(chk[:1000].cpu(), torch.from_numpy(src[:1000].astype(np.int32)))
