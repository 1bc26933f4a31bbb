"""Derived products of a sweep computed on the device so that scalars or small windows leave the GPU instead of N x N fp64
arrays: Strehl ratio (``docs/source/user/aberration/index.rst:27-45``; the reference documents it and has no code, so the
checker is numpy on the oracle's arrays), the on-axis / peak values, and the cropped / narrowed host product."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _focus_chain(pair_cls, n, coeff_nm):
    """A circular pupil with a few Zernike terms, focused by a lens: returns (oracle WFO, device WFO, oracle wfe)."""
    p = pair_cls(1.0, 1.0e-6, n, 4)
    p.call("aperture", 0.0, 0.0, r=0.5, shape="circular")
    p.call("make_stop")
    wfe = None
    if coeff_nm is not None:
        Z = np.array([0.0, 0.0, 0.0] + list(coeff_nm)) * 1e-9
        ro, _ = p.call("zernikes", np.arange(len(Z)), Z, "standard", True, 0.5)
        wfe = ro
    p.call("lens", 8.0)
    p.call("propagate", 8.0)
    return p, wfe


def test_strehl_ratio_and_marechal_against_numpy():
    import torch

    from helpers import Pair
    from paos_b200 import strehl

    n = 256
    ideal, _ = _focus_chain(Pair, n, None)
    aber, wfe = _focus_chain(Pair, n, [25.0, -15.0, 10.0, 5.0])
    ideal.check()
    aber.check()
    psf_i, psf_a = ideal.d.psf_device(), aber.d.psf_device()
    got = strehl.strehl_ratio(aber.d, psf_a, psf_i)
    ref_i, ref_a = np.abs(ideal.o._wfo) ** 2, np.abs(aber.o._wfo) ** 2
    want = ref_a[n // 2, n // 2] / ref_i[n // 2, n // 2]
    assert got == pytest.approx(want, rel=1e-9) and 0.5 < got < 1.0
    pk = strehl.psf_peak(aber.d, psf_a)
    aber.d.sync()
    pk = pk.cpu().numpy()
    assert pk[0] == pytest.approx(ref_a[n // 2, n // 2], rel=1e-9) and pk[1] == pytest.approx(ref_a.max(), rel=1e-9)
    # Marechal estimate from the wavefront-error screen over the pupil (rho <= 1)
    w = wfe.filled(0.0)
    x = (np.arange(n) - n // 2) * (4.0 / n)
    inside = (x[None, :] ** 2 + x[:, None] ** 2) / 0.25 <= 1.0
    var = np.mean(w[inside] ** 2) - np.mean(w[inside]) ** 2
    want_m = 1.0 - (2 * np.pi / 1.0e-6) ** 2 * var
    fresh = __import__("paos_b200").WFO(1.0, 1.0e-6, n, 4)
    got_m = strehl.strehl_marechal(fresh, torch.from_numpy(w).cuda(), 0.5)
    assert got_m == pytest.approx(want_m, rel=1e-10)
    assert got_m == pytest.approx(got, abs=0.03)  # the two definitions agree for a Strehl ratio this high
    got_np = strehl.strehl_marechal(fresh, wfe, 0.5)  # masked numpy input is uploaded
    assert got_np == pytest.approx(want_m, rel=1e-10)


@pytest.mark.parametrize("dtype,hdtype", [("complex128", "float32"), ("complex128", "float64"), ("complex64", "float32")])
def test_sweep_window_and_peaks(dtype, hdtype):
    import torch

    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    n, win = 256, 64
    jobs = configs.airs_ch0(grid=n, n_wl=7)
    sw = Sweep(n, slots=2, what="psf", batch=2, dtype=dtype)
    hd = getattr(torch, hdtype)
    host = torch.zeros((len(jobs), win, win), dtype=hd, pin_memory=True)
    peaks = torch.zeros((len(jobs), 2), dtype=torch.float64, device="cuda")
    x0 = y0 = (n - win) // 2
    out, meta = sw.run(jobs, host_out=host, host_window=(x0, y0 + 3, win, win), peak_out=peaks)
    full = out.cpu()
    want = full[:, y0 + 3: y0 + 3 + win, x0: x0 + win].to(hd)
    assert torch.equal(host, want)
    pk = peaks.cpu().numpy()
    assert np.array_equal(pk[:, 0], full[:, n // 2, n // 2].double().numpy())
    assert np.array_equal(pk[:, 1], full.reshape(len(jobs), -1).max(dim=1).values.double().numpy())
    with pytest.raises(ValueError):
        sw.run(jobs, host_out=host, host_window=(0, 0, win + 1, win))
    with pytest.raises(ValueError):
        sw.run(jobs, host_out=host, host_window=(n - 8, 0, win, win))  # window leaves the grid


def test_pipeline_default_call_saves_the_datacube(tmp_path, data_dir):
    """The reference's default call (save=True, store_keys amplitude,dx,dy,wl; pipeline.py:72-84, :157-172) must work:
    saved surfaces stream to pinned memory while the chain runs, then the data cube is written in the reference's layout."""
    import os

    import paos_b200
    from paos_b200.save_output import load_output

    conf = os.path.join(data_dir, "Hubble_simple.ini")
    out = str(tmp_path / "hubble.h5")
    pv = {"conf": conf, "output": out, "return": True}
    ret = paos_b200.pipeline(pv)
    assert len(ret) == 1 and os.path.isfile(pv["written"])
    pup, params, wls, fields, chains = paos_b200.parse_config(conf)
    ref = paos_b200.run(pup, 1e-6 * wls[0], params["grid_size"], params["zoom"], fields[0], chains[0])
    tree = load_output(pv["written"])
    group = tree[str(wls[0])]
    assert sorted(group) == [f"S{k:02d}" for k in sorted(ref)]
    for num, item in ref.items():
        g = group[f"S{num:02d}"]
        assert set(g) == {"amplitude", "dx", "dy", "wl"}
        assert np.array_equal(g["amplitude"], item["amplitude"]) and float(g["dx"]) == item["dx"] and float(g["wl"]) == item["wl"]
        # the async snapshots returned to the caller are the same arrays
        assert np.array_equal(ret[0][num]["amplitude"], item["amplitude"])
        assert np.array_equal(ret[0][num]["phase"], item["phase"]) and np.array_equal(ret[0][num]["wfo"], item["wfo"])
        assert ret[0][num]["propagator"] == item["propagator"]
    with pytest.raises(KeyError):
        paos_b200.pipeline({"conf": conf})  # save=True needs an output name, as in the reference
    with pytest.raises(NotImplementedError):
        paos_b200.pipeline({"conf": conf, "save": False, "plot": True})


def test_wfe_sweep_over_a_range_of_realizations(data_dir, tmp_path):
    """-wfe with a column range: every realization x wavelength through the batch front-end; realization c of
    the range equals the reference-style single-column pipeline call for that c."""
    import os
    import shutil

    import paos_b200

    conf = str(tmp_path / "fgs1.ini")
    text = open(os.path.join(data_dir, "Ariel_FGS-FGS1.ini")).read()
    lines, section = [], None
    for ln in text.splitlines():
        if ln.startswith("["):
            section = ln.strip()
        if section == "[general]" and ln.startswith("grid_size"):
            ln = "grid_size = 256"
        if section == "[lens_13]" and ln.startswith("ignore"):
            ln = "ignore = False"
        lines.append(ln)
    open(conf, "w").write("\n".join(lines) + "\n")
    wfe = os.path.join(data_dir, "wfe_realization_SN20210914.csv")
    stack, meta, index = paos_b200.wfe_sweep({"conf": conf, "wfe": f"{wfe},2-4"})
    n_wl = len(paos_b200.parse_config(conf)[2])
    assert stack.shape[0] == 3 * n_wl and [c for c, _ in index][::n_wl] == [2, 3, 4]
    single = paos_b200.pipeline({"conf": conf, "wfe": f"{wfe},3", "light_output": True, "save": False, "return": True})
    for k in range(n_wl):
        ret = single[k]
        ref = ret[max(ret)]["amplitude"] ** 2
        got = stack[n_wl + k].cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-12 * ref.max()
    assert not np.array_equal(stack[0].cpu().numpy(), stack[n_wl].cpu().numpy())
