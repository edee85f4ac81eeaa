"""How expensive is page-locking host memory on this box, and what does a pageable copy cost instead?
Decides how fastore_bin_b200 (the CLI) should hold its chunk buffers: numbers land in profiles/ and DESIGN.md section 8."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

GB = 1 << 30


def t(f):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, r


def main():
    size = int(float(sys.argv[1]) * GB) if len(sys.argv) > 1 else GB
    t0 = time.perf_counter()
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    out = {"bytes": size, "context_s": time.perf_counter() - t0}
    rt = torch.cuda.cudart()
    dev = torch.empty(size, dtype=torch.uint8, device="cuda")
    # 1) cudaHostAlloc through torch's pinned allocator
    dt, pinned = t(lambda: torch.empty(size, dtype=torch.uint8, pin_memory=True))
    out["host_alloc_pinned_s"] = dt
    dt, _ = t(lambda: dev.copy_(pinned, non_blocking=True))
    out["h2d_pinned_first_GBps"] = size / dt / 1e9
    dt, _ = t(lambda: dev.copy_(pinned, non_blocking=True))
    out["h2d_pinned_GBps"] = size / dt / 1e9
    dt, pinned2 = t(lambda: torch.empty(size, dtype=torch.uint8, pin_memory=True))
    out["host_alloc_pinned_second_s"] = dt
    del pinned2
    # 2) pageable memory: first touch, then copies
    t0 = time.perf_counter()
    page = np.empty(size, dtype=np.uint8)
    page[::4096] = 1
    out["pageable_first_touch_s"] = time.perf_counter() - t0
    pt = torch.from_numpy(page)
    dt, _ = t(lambda: dev.copy_(pt))
    out["h2d_pageable_first_GBps"] = size / dt / 1e9
    dt, _ = t(lambda: dev.copy_(pt))
    out["h2d_pageable_GBps"] = size / dt / 1e9
    dt, _ = t(lambda: pt.copy_(dev))
    out["d2h_pageable_GBps"] = size / dt / 1e9
    # 3) cudaHostRegister of the touched pageable buffer
    dt, rc = t(lambda: rt.cudaHostRegister(page.ctypes.data, size, 0))
    out["host_register_s"] = dt
    out["host_register_rc"] = int(rc)
    dt, _ = t(lambda: dev.copy_(pt, non_blocking=True))
    out["h2d_registered_GBps"] = size / dt / 1e9
    dt, _ = t(lambda: rt.cudaHostUnregister(page.ctypes.data))
    out["host_unregister_s"] = dt
    # 4) plain memcpy speed of the host (one thread), for scale
    a = np.empty(size, dtype=np.uint8)
    a[::4096] = 1
    t0 = time.perf_counter()
    np.copyto(a, page)
    out["host_memcpy_GBps"] = size / (time.perf_counter() - t0) / 1e9
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
