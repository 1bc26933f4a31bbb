#!/usr/bin/env python
"""Device time of paos_encircled_energy alone (histogram + scan) on a resident 2048^2 PSF."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import paos_b200
from paos_b200 import ee

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
w = paos_b200.WFO(1.0, 1e-6, n, 4)
w.aperture(0.0, 0.0, r=0.5, shape="circular")
w.make_stop()
w.lens(10.0)
w.propagate(10.0)
psf = w.psf_device()
w.sync()
for nb in (64, 256, 4096):
    out = torch.empty(nb + 1, dtype=torch.float64, device="cuda")
    for _ in range(10):
        ee.encircled_energy(w, psf, w.dx, w.dy, w.fratio, w.wl, 8.0, nb, out=out)
    w.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(w._stream)
    for _ in range(100):
        ee.encircled_energy(w, psf, w.dx, w.dy, w.fratio, w.wl, 8.0, nb, out=out)
    b.record(w._stream)
    w.sync()
    print(f"n={n} nbins={nb}: {a.elapsed_time(b) * 10:.1f} us per call")
