#!/usr/bin/env python
"""Micro-benchmark of the line-pass kernel: K chained shifted FFT2s become ONE row pass and ONE column pass of
K line FFTs each.  Prints per-kind average device time (CUDA events around every launch, one stream).

    python tools/kernel_bench.py --n 2048 --k 1 2 4 8 [--dtype complex64] [--reps 20]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--k", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--reps", type=int, default=60)
    ap.add_argument("--warm", type=int, default=60)
    ap.add_argument("--fields", type=int, default=4, help="distinct wavefronts cycled through (defeats L2 reuse)")
    ap.add_argument("--gen", action="store_true", help="put an elliptical mask between the FFT2s")
    args = ap.parse_args()
    import paos_b200
    from paos_b200 import _lib

    n = args.n
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    import torch

    stream = torch.cuda.Stream()  # ONE explicit stream for all wavefronts: launches never overlap
    ws = []
    for _ in range(args.fields):
        w = paos_b200.WFO(1.0, 1e-6, n, 4, dtype=args.dtype, stream=stream)
        w.wfo = x
        _lib.check(_lib.lib.paos_wfo_enable_timing(w._handle, 1))
        ws.append(w)
    elem = 16 if args.dtype == "complex128" else 8
    sweep_bytes = 2 * elem * n * n
    print(f"n={n} dtype={args.dtype} sweep={sweep_bytes/1e6:.1f} MB (read+write)")
    for k in args.k:
        for rep in range(args.reps + args.warm):
            w = ws[rep % len(ws)]
            for i in range(k):
                w._fft2(inverse=bool(i & 1))
                if args.gen and i + 1 < k:
                    w.aperture(0.0, 0.0, r=0.7, shape="circular")
            w.flush()
            if rep == args.warm - 1:  # warm-up done (clocks ramped): clear the buckets
                for ww in ws:
                    ww.sync()
                    for col in (0, 1):
                        for nf in range(_lib.MAX_CHAINED_FFTS + 1):
                            _lib.lib.paos_wfo_timing_detail(ww._handle, col, nf, None, None, 1)
        out = {}
        for ww in ws:
            ww.sync()
            for col in (0, 1):
                for nf in range(_lib.MAX_CHAINED_FFTS + 1):
                    ms, cnt = C.c_double(), C.c_uint64()
                    _lib.lib.paos_wfo_timing_detail(ww._handle, col, nf, C.byref(ms), C.byref(cnt), 1)
                    if cnt.value:
                        a = out.get((col, nf), [0.0, 0])
                        a[0] += ms.value
                        a[1] += cnt.value
                        out[(col, nf)] = a
        import subprocess
        try:
            clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        except OSError:
            clk = "?"
        print(f"  [clocks right after K={k}: {clk}]")
        for (col, nf), (ms, cnt) in sorted(out.items()):
            us = 1e3 * ms / cnt
            print(f"  K={k} {'col' if col else 'row'}x{nf}: {us:8.1f} us/launch  {sweep_bytes/us/1e3:7.0f} GB/s sweep  "
                  f"{nf*sweep_bytes/us/1e3:7.0f} GB/s algorithmic  ({cnt} launches)")


if __name__ == "__main__":
    main()
