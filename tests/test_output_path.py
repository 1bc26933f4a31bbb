"""Output path (``paos_b200.save_output``: reference ``paos/core/saveOutput.py:120-303``) and the ``paos`` import surface
(reference ``paos/__init__.py:39-48``).  CPU-only: the writers work on result dictionaries, whoever produced them."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Aperture:
    def __init__(self):
        self.positions = np.array([32.0, 32.0])
        self.a, self.b, self.theta = 10.0, 8.0, 0.0


def _retval(seed, with_wfe_on_second=True):
    from paos_b200.abcd import ABCD

    rng = np.random.default_rng(seed)
    out = {}
    for num in (2, 7):
        item = dict(aperture=_Aperture() if num == 2 else None, amplitude=rng.standard_normal((8, 8)), phase=rng.standard_normal((8, 8)),
                    wfo=rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8)), wz=0.5, distancetofocus=-1.0, fratio=np.inf,
                    dx=1e-3, dy=2e-3, wl=1e-6, extent=(-1.0, 1.0, -2.0, 2.0), propagator="II", ABCDt=ABCD(thickness=1.0), ABCDs=ABCD())
        if num == 7 and with_wfe_on_second:
            item["wfe"] = np.ma.MaskedArray(rng.standard_normal((8, 8)), mask=rng.random((8, 8)) > 0.5, fill_value=0.0)
        out[num] = item
    return out


def test_save_output_layout_and_round_trip(tmp_path):
    from paos_b200.save_output import have_h5py, load_output, save_output

    ret = _retval(1)
    path = save_output(ret, str(tmp_path / "one.h5"), keys_to_keep=["amplitude", "dx", "dy", "wl", "aperture", "ABCDt"])
    assert os.path.isfile(path) and path.endswith(".h5" if have_h5py() else ".npz")
    tree = load_output(path)
    assert set(tree) == {"info", "S02", "S07"}
    assert {"file_name", "file_time", "creator", "program_name", "program_version"} <= set(tree["info"])
    assert set(tree["S02"]) == {"amplitude", "dx", "dy", "wl", "aperture", "ABCDt"}
    assert np.array_equal(tree["S02"]["amplitude"], ret[2]["amplitude"]) and float(tree["S07"]["dy"]) == 2e-3
    assert np.array_equal(tree["S02"]["aperture"]["positions"], [32.0, 32.0]) and float(tree["S02"]["aperture"]["a"]) == 10.0
    assert "aperture" not in tree["S07"]  # None is skipped with a warning (saveOutput.py:76-78)
    assert np.array_equal(tree["S07"]["ABCDt"]["_ABCD"], ret[7]["ABCDt"]())


def test_keys_of_the_first_surface_decide_when_no_list_is_given(tmp_path):
    """saveOutput.py:151-152: keys_to_keep is taken from the first surface and then applies to all of them."""
    from paos_b200.save_output import load_output, save_output

    tree = load_output(save_output(_retval(2), str(tmp_path / "all.npz")))
    assert "wfe" not in tree["S07"] and "phase" in tree["S07"] and tree["S07"]["wfo"].dtype == np.complex128
    assert str(tree["S02"]["propagator"]) == "II" and tuple(tree["S02"]["extent"]) == (-1.0, 1.0, -2.0, 2.0)


def test_datacube_groups_and_overwrite(tmp_path):
    from paos_b200.save_output import load_output, save_datacube

    rets = [_retval(3), _retval(4)]
    name = str(tmp_path / "cube.npz")
    save_datacube(rets, name, ["1.95", "3.9"], keys_to_keep=["amplitude", "wl"])
    tree = load_output(name)
    assert set(tree) == {"info", "1.95", "3.9"} and set(tree["3.9"]) == {"S02", "S07"}
    assert np.array_equal(tree["3.9"]["S07"]["amplitude"], rets[1][7]["amplitude"])
    save_datacube(rets[:1], name, ["only"], keys_to_keep=["wl"], overwrite=True)
    assert set(load_output(name)) == {"info", "only"}
    with pytest.raises(AssertionError):
        save_datacube(rets, name, "not a list")
    with pytest.raises(NameError):
        bad = _retval(5)
        bad[2]["dx"] = object()
        save_datacube([bad], name, ["x"])


def test_paos_import_surface_resolves_to_the_device_path():
    """`import paos; paos.core.run.run` etc. (reference paos/__init__.py:39-48) -- in a fresh interpreter, because the
    oracle's stub loader registers the reference under the same top-level name in this one."""
    code = """
import paos
from paos.core.run import run, push_results
from paos.core.parseConfig import parse_config
from paos.core.pipeline import pipeline
from paos.core.saveOutput import save_datacube, save_output
from paos.core.coordinateBreak import coordinate_break
from paos.core.raytrace import raytrace
from paos.classes.wfo import WFO
from paos.classes.abcd import ABCD
from paos.classes.zernike import Zernike, PolyOrthoNorm
from paos.classes.psd import PSD
from paos.util.material import Material
import paos_b200
assert paos.run is run is paos_b200.run and paos.WFO is WFO is paos_b200.WFO
assert paos.core.run.run is run and paos.classes.wfo.WFO is WFO
for name in ("ABCD", "PSD", "WFO", "Zernike", "PolyOrthoNorm", "coordinate_break", "parse_config", "plot_pop", "raytrace", "run",
             "save_datacube", "save_output"):
    assert hasattr(paos, name), name
try:
    paos.plot_pop({})
except NotImplementedError:
    print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
