#!/usr/bin/env python
"""Per-kernel-group timing of the device multinomial resampler (smc_multinomial_ancestors + gather) at N = 2^k."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import torch
    from particle_filters_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda", 0)
    for k in (20, 24, 26):
        n = 1 << k
        g = torch.Generator(device=dev).manual_seed(1)
        w = torch.rand((n,), dtype=torch.float64, device=dev, generator=g) ** 4
        w /= w.sum()
        x = torch.randn((n,), dtype=torch.float32, device=dev)
        y = torch.empty_like(x)
        anc = torch.empty((n,), dtype=torch.int32, device=dev)
        nb = int(lib.smc_multinomial_workspace_bytes(n))
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

        def draw(t):
            _lib.check(lib.smc_multinomial_ancestors(C.c_void_p(w.data_ptr()), n, n, None, C.c_uint64(7), t,
                                                     C.c_void_p(anc.data_ptr()), C.c_void_p(ws.data_ptr()),
                                                     C.c_uint64(nb), st))

        def gather():
            _lib.check(lib.smc_gather(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(anc.data_ptr()),
                                      n, n, 1, _lib.SMC_F32, st))

        out = {"n": n}
        for name, fn in (("ancestors_ms", draw), ("gather_ms", lambda t: gather())):
            for t in range(3):
                fn(t)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for t in range(10):
                fn(3 + t)
            e1.record()
            torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) / 10
        out["particles_per_s"] = n / ((out["ancestors_ms"] + out["gather_ms"]) * 1e-3)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
