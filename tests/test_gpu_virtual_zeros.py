"""GPU tests of the "virtual zeros" bookkeeping: lines blanked by an aperture are neither loaded nor stored, the planner
remembers the zero band and every other reader of the field sees real zeros."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import Pair, TOL, random_field, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [128, 512])
def test_every_reader_sees_the_zeros(n):
    """aperture -> flush leaves stale memory outside the aperture's rows; the complex copy, the stand-alone read-outs and
    the borrowed tensor must all show zeros there, also when the stale content was random data."""
    p = Pair(1.0, 1e-6, n, 4, field=random_field(n, 3))
    p.call("aperture", 0.03, -0.02, hx=0.21, hy=0.13, shape="elliptical")
    p.d.flush()
    st = p.d.stats()
    amp = p.d.amplitude          # stand-alone read-out kernel on a field with a band
    assert relerr(amp, p.o.amplitude) <= 1e-12
    assert relerr(p.d.wfo, p.o._wfo) <= 1e-12
    t = p.d.field_tensor()
    p.d.sync()
    assert relerr(t.cpu().numpy(), p.o._wfo) <= 1e-12
    assert p.d.stats()["kernel_launches"] > st["kernel_launches"]
    p.check()


def test_band_handed_from_pass_to_pass():
    """rows blanked -> column pass loads only the band -> columns blanked -> row pass; reads in between and at the end."""
    n = 256
    p = Pair(1.0, 2e-6, n, 4, field=random_field(n, 5))
    p.call("aperture", 0.0, 0.05, hx=0.3, hy=0.1, shape="elliptical")
    p.call("lens", 3.0)
    p.call("propagate", 1.5)
    p.check()
    d = p.o.dx
    p.call("aperture", -3 * d, 0.0, hx=20 * d, hy=60 * d, shape="elliptical")
    p.call("propagate", 0.7)
    d = p.o.dx
    p.call("aperture", 0.0, 0.0, hx=80 * d, hy=50 * d, shape="rectangular")
    p.call("make_stop")
    p.call("propagate", 0.8)
    p.check()


def test_stop_reduction_on_a_banded_field():
    n = 256
    p = Pair(1.0, 2e-6, n, 4, field=random_field(n, 7))
    p.call("aperture", 0.0, 0.0, hx=0.2, hy=0.3, shape="elliptical")
    p.d.flush()               # band left in memory
    p.call("make_stop")       # reduction reads the stored field
    p.check()
    assert abs(np.sum(np.abs(p.d.wfo) ** 2) - 1.0) <= 1e-12


def test_aperture_that_misses_the_grid():
    """an aperture whose bounding box misses the grid blanks everything (empty band)"""
    n = 128
    p = Pair(1.0, 1e-6, n, 1, field=random_field(n, 9))
    p.call("aperture", 5.0, 0.0, hx=0.1, hy=0.1, shape="elliptical")
    p.call("lens", 2.0)
    p.call("propagate", 1.0)
    assert np.all(p.o._wfo == 0)
    assert np.all(p.d.wfo == 0) and np.all(p.d.amplitude == 0)


_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
from paos_b200 import configs
from paos_b200.sweep import Sweep
jobs = configs.airs_ch0(grid=256, n_wl=3) + configs.hubble(grid=256)
sw = Sweep(256, slots=1, what="amplitude")
out, _ = sw.run(jobs)
np.save({path!r}, out.cpu().numpy())
"""


def test_zero_fill_mode_gives_identical_results(tmp_path):
    """PAOS_ZERO_FILL=1 (zeros really stored, no band kept) and the default (virtual zeros) must agree bit for bit."""
    outs = []
    for mode in ("0", "1"):
        path = str(tmp_path / f"amp_{mode}.npy")
        env = dict(os.environ)
        env.pop("PAOS_ZERO_FILL", None)
        if mode == "1":
            env["PAOS_ZERO_FILL"] = "1"
        subprocess.run([sys.executable, "-c", _SCRIPT.format(root=ROOT, path=path)], check=True, env=env, timeout=300)
        outs.append(np.load(path))
    assert outs[0].shape == outs[1].shape and np.array_equal(outs[0], outs[1])


def test_final_read_out_drops_the_field_and_says_so():
    """paos_wfo_read_device_final: same numbers as the ordinary read-out; afterwards the handle refuses work until it is
    reset."""
    import ctypes as C

    import torch

    import paos_b200
    from paos_b200 import _lib

    n = 256
    res = []
    for final in (False, True):
        w = paos_b200.WFO(1.0, 2e-6, n, 4)
        w.aperture(0.0, 0.0, hx=0.3, hy=0.2, shape="elliptical")
        w.make_stop()
        w.lens(2.0)
        w.propagate(2.0)
        out = torch.empty((n, n), dtype=torch.float64, device="cuda")
        fn = _lib.lib.paos_wfo_read_device_final if final else _lib.lib.paos_wfo_read_device
        _lib.check(fn(w._handle, _lib.READ_PSF, C.c_void_p(out.data_ptr())))
        w.sync()
        res.append(out.cpu().numpy())
        if final:
            with pytest.raises(paos_b200.PaosError):
                w.amplitude
            w.lens(1.0)
            with pytest.raises(paos_b200.PaosError):
                w.flush()
            _lib.check(_lib.lib.paos_wfo_reset(w._handle))
            assert np.all(w.amplitude == 1.0)
    assert np.array_equal(res[0], res[1]) and res[0].sum() > 0.99
    with pytest.raises(ValueError):
        w2 = paos_b200.WFO(1.0, 2e-6, n, 4)
        _lib.check(_lib.lib.paos_wfo_read_device_final(w2._handle, _lib.READ_WFO, C.c_void_p(out.data_ptr())))


def test_sweep_slot_is_reusable_after_final_read_outs():
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    jobs = configs.airs_ch0(grid=256, n_wl=4)
    sw = Sweep(256, slots=1, what="psf")
    a, _ = sw.run(jobs)
    a = a.cpu().numpy().copy()
    b, _ = sw.run(jobs)  # same slots again: every chain starts with a reset
    assert np.array_equal(a, b.cpu().numpy())
    energy = a.sum(axis=(1, 2))  # normalised at the stop, then clipped by the apertures behind it
    assert np.all(energy > 0.1) and np.all(energy <= 1.0 + 1e-9), energy


def test_edge_tables_do_not_change_a_bit(tmp_path):
    """The rim pixels of an elliptical mask take their exact overlap from a table built once per pass (one lane per pixel)
    instead of from the routine called by the lane that meets them; the factor is the same double either way, so every
    output must be identical with PAOS_NO_EDGE_TABLES=1 (read once per process: the comparison runs in two children)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
import paos_b200
from paos_b200 import configs
from paos_b200.sweep import Sweep
out = []
jobs = configs.airs_ch0(grid=512, n_wl=5) + configs.airs_ch0(grid=512, n_wl=3, wl_range=(0.8, 1.2))
o, _ = Sweep(512, slots=1, what="amplitude", batch=4).run(jobs)
out.append(o.cpu().numpy())
w = paos_b200.WFO(1.0, 1e-6, 512, 4)
for (xc, yc, hx, hy, ob) in [(0.0, 0.0, 0.5, 0.3, False), (0.05, -0.02, 0.1, 0.12, True), (0.3, 0.2, 0.02, 0.01, True), (1.9, 0.0, 0.4, 0.4, True)]:
    w.aperture(xc, yc, hx=hx, hy=hy, obscuration=ob)
    w._fft2()
out.append(np.abs(w.wfo))
job = configs.hubble(grid=256)[0]
r = paos_b200.run(job["pupil_diameter"], job["wavelength"], 256, job["zoom"], job["field"], job["opt_chain"])
out += [r[k]["amplitude"] for k in sorted(r)]
np.savez(sys.argv[1], *out)
""" % root
    files = []
    for tag, env in (("tables", {}), ("inline", {"PAOS_NO_EDGE_TABLES": "1"})):
        path = str(tmp_path / f"{tag}.npz")
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code, path], capture_output=True, text=True, env=e)
        assert r.returncode == 0, r.stderr[-2000:]
        files.append(np.load(path))
    a, b = files
    assert sorted(a.files) == sorted(b.files) and len(a.files) >= 5
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
