"""CPU tests of the host-side logic: paraxial scalars, job builders, sharding, and the N > 1 gather on gloo."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abcd_matches_oracle():
    import paos_b200
    from oracle import paos_np

    rng = np.random.default_rng(0)
    for _ in range(50):
        t, c, n1, n2, M = rng.uniform(0.1, 3), rng.uniform(-2, 2), rng.choice([1.0, -1.0, 1.5]), rng.choice([1.0, -1.0, 1.43]), rng.uniform(0.5, 2)
        a, b = paos_b200.ABCD(t, c, n1, n2, M), paos_np.ABCD(t, c, n1, n2, M)
        assert np.array_equal(a(), b())
        for p in ("thickness", "M", "n1n2", "power", "f_eff", "cin", "cout"):
            assert getattr(a, p) == getattr(b, p)
        prod_a, prod_b = a * paos_b200.ABCD(0.3, 0.1), b * paos_np.ABCD(0.3, 0.1)
        assert np.array_equal(prod_a(), prod_b()) and prod_a.cout == prod_b.cout
    with pytest.raises(ValueError):
        paos_b200.ABCD(n1=0.0)


def test_coordinate_break_matches_oracle():
    import paos_b200
    from oracle import paos_np

    rng = np.random.default_rng(1)
    for _ in range(20):
        vt, vs = rng.normal(size=2) * 0.01, rng.normal(size=2) * 0.01
        args = (rng.normal() * 0.01, rng.normal() * 0.01, rng.normal() * 3, rng.normal() * 3, 0.0)
        a = paos_b200.coordinate_break(vt, vs, *args)
        b = paos_np.coordinate_break(vt, vs, *args)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    a = paos_b200.coordinate_break(np.array([0.0, 0.1]), np.array([0.0, 0.0]), np.nan, np.nan, np.nan, 2.0, 0.0)
    b = paos_np.coordinate_break(np.array([0.0, 0.1]), np.array([0.0, 0.0]), np.nan, np.nan, np.nan, 2.0, 0.0)
    assert np.array_equal(a[0], b[0])
    with pytest.raises(ValueError):
        paos_b200.coordinate_break(np.zeros(2), np.zeros(2), 0, 0, 0, 0, 0, order=1)


def test_config_builders():
    from paos_b200 import configs

    jobs = configs.airs_ch0(grid=2048, n_wl=256)
    assert len(jobs) == 256 and jobs[0]["gridsize"] == 2048
    assert abs(jobs[0]["wavelength"] - 1.95e-6) < 1e-18 and abs(jobs[-1]["wavelength"] - 3.9e-6) < 1e-18
    saved = [it["name"] for it in jobs[0]["opt_chain"].values() if it["save"]]
    assert saved == ["IMAGE_PLANE"]
    assert jobs[3]["opt_chain"] is not jobs[4]["opt_chain"]
    h = configs.hubble()[0]
    assert h["gridsize"] == 1024 and h["zoom"] == 4 and abs(h["wavelength"] - 1e-6) < 1e-20
    f = configs.fgs1_montecarlo(grid=512, realizations=[0, 999])
    z0, z1 = f[0]["opt_chain"][13]["Z"], f[1]["opt_chain"][13]["Z"]
    assert len(z0) == 36 and np.all(z0[:3] == 0) and not np.array_equal(z0, z1)
    t = configs.ta_ground_psd(grid=1024, n_wl=4)
    assert len(t) == 36 and len({j["psd_seed"] for j in t}) == 36
    assert any(it["type"] == "PSD" for it in t[0]["opt_chain"].values())
    n1, n2 = configs.psd_noise_from_seed(5)(0, (4, 4))
    rs = np.random.RandomState(5)
    assert np.array_equal(n1, rs.randn(4, 4)) and np.array_equal(n2, rs.randn(4, 4))


def test_grid_sag_config_is_on_grid(tmp_path):
    from paos_b200 import configs

    job = configs.grid_sag(grid=256, wavelengths=(0.55, 3.0), workdir=str(tmp_path))[1]
    sag = [it for it in job["opt_chain"].values() if it["type"] == "Grid Sag"][0]
    d = job["pupil_diameter"] * job["zoom"] / 256
    assert sag["nx"] == 256 and sag["delx"] == d and sag["grid_sag"].shape == (256, 256)
    assert np.max(np.abs(sag["grid_sag"])) <= 30e-9 + 1e-20


def test_partition_is_contiguous_and_balanced():
    from paos_b200 import configs
    from paos_b200.sweep import job_cost, partition

    jobs = configs.airs_ch0(grid=64, n_wl=37)
    for world in (1, 2, 3, 8):
        blocks = partition(jobs, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == len(jobs)
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 2
    assert job_cost(jobs[0]) >= 10
    assert partition(jobs[:2], 4) == [(0, 0), (0, 1), (1, 1), (1, 2)] or sum(b - a for a, b in partition(jobs[:2], 4)) == 2


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from paos_b200.sweep import gather_stack, partition

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        jobs = [{"opt_chain": {}} for _ in range(5)]  # 5 jobs over 2 ranks: ragged blocks (3 + 2 or 2 + 3)
        blocks = partition(jobs, world)
        lo, hi = blocks[rank]
        counts = [b - a for a, b in blocks]
        local = torch.stack([torch.full((4, 4), float(k), dtype=torch.float64) for k in range(lo, hi)]) if hi > lo else torch.zeros((0, 4, 4), dtype=torch.float64)
        full = gather_stack(local, counts, dst=0)
        if rank == 0:
            assert full.shape == (5, 4, 4)
            assert [float(full[k, 0, 0]) for k in range(5)] == [0.0, 1.0, 2.0, 3.0, 4.0]
            open(os.path.join(tmp, "ok"), "w").write("ok")
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts,chunk", [([4, 4], 2), ([5, 3], 2), ([3, 0, 7], 4), ([1, 1, 1, 1], 8), ([6], 4)])
def test_chunk_schedule_covers_every_row_once(counts, chunk):
    """The overlapped gather (ChunkGather) sends chunk g of every rank to its final rows: together the chunks must place
    every PSF exactly once, in rank order, for equal and ragged blocks."""
    from paos_b200.sweep import chunk_schedule

    rng = np.random.default_rng(1)
    locals_ = [rng.standard_normal((c, 3, 3)) for c in counts]
    if all(c == 0 for c in counts):
        return
    ref = [a for a in locals_ if len(a)]
    row = ref[0][0].nbytes
    locals_ = [a if len(a) else np.zeros((0, 3, 3)) for a in locals_]
    full = np.zeros(sum(counts) * row, dtype=np.uint8)
    covered = np.zeros(sum(counts), dtype=int)
    for lo, sizes, offs in chunk_schedule(counts, chunk, row):
        assert len(sizes) == len(offs) == len(counts)
        for q, loc in enumerate(locals_):
            if sizes[q]:
                assert sizes[q] % row == 0 and offs[q] % row == 0
                full[offs[q]: offs[q] + sizes[q]] = loc.view(np.uint8).reshape(-1)[lo * row: lo * row + sizes[q]]
                covered[offs[q] // row: (offs[q] + sizes[q]) // row] += 1
    assert np.all(covered == 1)
    assert np.array_equal(full.view(np.float64).reshape(-1, 3, 3), np.concatenate(locals_, axis=0))


def _gloo_schedule_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    from paos_b200.sweep import chunk_schedule, partition

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        jobs = [{"opt_chain": {}} for _ in range(11)]
        blocks = partition(jobs, world)
        counts = [b - a for a, b in blocks]
        sched = chunk_schedule(counts, 4, 128)
        # every rank must issue the same sequence of collective calls: compare the schedules
        mine = torch.tensor([v for lo, sizes, offs in sched for v in [lo] + sizes + offs], dtype=torch.int64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        assert all(torch.equal(g, mine) for g in gathered)
        if rank == 0:
            open(os.path.join(tmp, "ok2"), "w").write("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_chunk_schedule_agrees_across_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = 29900 + os.getpid() % 90
    mp.spawn(_gloo_schedule_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok2")


def test_gather_stack_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = 29600 + os.getpid() % 300
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")


@pytest.mark.parametrize("ny,nx,pitch,xdec,ydec", [
    (64, 64, (0.7, 0.7), 0.0, 0.0), (40, 48, (1.9, 1.6), 0.0, 0.0), (65, 64, (1.0, 1.0), 0.0, 0.0),
    (51, 77, (1.3, 0.8), -0.7, 2.2), (838 // 4, 1158 // 4, (0.37, 0.37), 0.0, 0.0),
])
def test_grid_sag_resampling_matches_oracle(ny, nx, pitch, xdec, ydec):
    """oracle.sag_host.prepare_sag (the separable operators the device kernels run, stated on the host) against the oracle's grid_sag (scipy.ndimage restatement
    of the skimage calls, wfo.py:696-862): same mask, same screen to rounding."""
    from oracle import paos_np
    from oracle.sag_host import prepare_sag

    n = 64
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:ny, 0:nx]
    sag = 30e-9 * np.cos(2 * np.pi * xx / 17.0) * np.sin(2 * np.pi * yy / 13.0) + rng.standard_normal((ny, nx)) * 1e-9
    sag[:3, :] = 0.0
    sag[5, 7] = np.nan
    o = paos_np.WFO(1.0, 1e-6, n, 2)
    ro = o.grid_sag(sag.copy(), nx, ny, pitch[0] * o.dx, pitch[1] * o.dy, xdec, ydec)
    screen, mask = prepare_sag(sag.copy(), nx, ny, pitch[0] * o.dx, pitch[1] * o.dy, xdec, ydec, n, o.dx, o.dy)
    assert screen.shape == (n, n) and np.array_equal(mask, ro.mask)
    assert np.max(np.abs(screen - ro.filled(0))) <= 1e-12 * np.max(np.abs(ro.filled(0)))


def test_resampler_known_answers():
    """Properties of the cubic resampler that hold whatever library restates it: constants and linear ramps inside the
    clip range are reproduced, identity scale returns the input, shapes follow round(scale * shape)."""
    from oracle import separable_resample as resample

    a = np.full((9, 14), 2.5)
    assert np.allclose(resample.rescale(a, (1.7, 0.6), True), 2.5, atol=1e-14)
    rng = np.random.default_rng(0)
    b = rng.standard_normal((12, 10))
    assert np.max(np.abs(resample.rescale(b, (1.0, 1.0), False) - b)) <= 1e-14
    assert resample.rescale(b, (2.0, 0.5), True).shape == (24, 5)
    assert resample.rescale(b, (0.01, 0.01), True).shape == (1, 1)
    ramp = np.add.outer(np.arange(60.0), 2.0 * np.arange(50.0))
    up = resample.rescale(ramp, (2.0, 2.0), False)
    yy = (np.arange(120) + 0.5) / 2 - 0.5
    xx = (np.arange(100) + 0.5) / 2 - 0.5
    want = np.add.outer(yy, 2.0 * xx)
    # a cubic spline reproduces a ramp exactly; the kink of the mirrored border decays as 0.268^k (k input pixels)
    inner = (slice(36, -36), slice(36, -36))
    assert np.max(np.abs(up[inner] - want[inner])) <= 1e-7


def test_oracle_encircled_energy_of_a_gaussian():
    """The oracle's discrete encircled energy against the closed form for a Gaussian spot: 1 - exp(-R^2 / (2 sigma^2))."""
    from oracle import paos_np

    n, dx, sigma = 512, 0.01, 0.2
    x = (np.arange(n) - n / 2) * dx
    psf = np.exp(-(x[None, :] ** 2 + x[:, None] ** 2) / (2 * sigma**2))
    ee, total = paos_np.encircled_energy(psf, dx, dx, 1.0, 1.0, 100)
    R = (np.arange(100) + 1) / 100.0
    assert total == pytest.approx(2 * np.pi * sigma**2 / dx**2, rel=1e-6)
    assert np.max(np.abs(ee - (1 - np.exp(-R**2 / (2 * sigma**2))))) < 8e-3  # pixel-centre binning at dx = sigma/20
    assert np.all(np.diff(ee) >= 0)


def test_native_compile_takes_grid_sag_anywhere_in_the_chain(tmp_path):
    """The native runner carries the raw map and lets the library resample it at the pitch of the surface (wfo.py:848-862),
    so a Grid Sag surface behind a propagation compiles like one at the INIT pitch; maps that are the same object share one
    content key (one prepared screen per sweep)."""
    import copy

    from paos_b200 import chain as chain_mod
    from paos_b200 import configs

    job = configs.grid_sag(grid=64, wavelengths=(3.0,), workdir=str(tmp_path))[0]
    cache = {}
    cc = chain_mod.compile_job(job, screen_cache=cache)
    rec = cc.array[1]
    assert rec.type == chain_mod.SURF_GRIDSAG and rec.sag_nx == 64 and rec.sag_ny == 64 and rec.sag_key != 0
    assert rec.sag_delx == job["pupil_diameter"] * job["zoom"] / 64
    late = copy.copy(job["opt_chain"][3])
    late.update(num=5.5, name="Sag2")
    moved = {k: v for k, v in job["opt_chain"].items()}
    moved[5.5] = late
    job2 = dict(job, opt_chain=dict(sorted(moved.items())))
    job2.pop("_compiled", None)
    cc2 = chain_mod.compile_job(job2, screen_cache=cache)
    keys = [r.sag_key for r in cc2.array[: cc2.count] if r.type == chain_mod.SURF_GRIDSAG]
    assert len(keys) == 2 and keys[0] == keys[1] == rec.sag_key


def test_grid_sag_refuses_absurd_padding():
    from oracle.sag_host import prepare_sag

    with pytest.raises(ValueError):
        prepare_sag(np.ones((90, 70)), 70, 90, 6e-5, 6e-5, 0.0, 0.0, 256, 0.0172, 0.0172)


def test_pipeline_chain_setup_and_refusals():
    """paos_b200.pipeline's option handling (pipeline.py:88-129): light_output keeps only IMAGE_PLANE, the -wfe option
    overrides Z1 with column c + 4 of the realization table; the default save=True needs an output name (as in the
    reference, which indexes passvalue['output']), plot requests are refused, not dropped."""
    import paos_b200
    from paos_b200 import configs

    pl = sys.modules["paos_b200.pipeline"]  # the package re-exports the function under the module's name, like the reference
    conf = os.path.join(ROOT, "paos_b200", "lens_data", "Ariel_FGS-FGS1.ini")
    csv = os.path.join(ROOT, "paos_b200", "lens_data", "wfe_realization_SN20210914.csv")
    pup, params, wls, field, chains = pl.setup_chains({"conf": conf, "light_output": True, "wfe": f"{csv},7"})
    assert len(chains) == len(wls) and field == paos_b200.parse_config(conf)[3][0]
    for chain in chains:
        saved = [it["name"] for it in chain.values() if it["save"]]
        assert saved == ["IMAGE_PLANE"]
        for it in chain.values():  # Z1 is 'ignore = True' in the shipped file: the override is dead unless it is enabled
            assert it["name"] != "Z1"
    want = np.append(np.zeros(3), configs.wfe_table()[:, 3 + 7] * 1e-9)
    assert np.array_equal(np.append(np.zeros(3), pl.read_wfe_column(csv, "7")), want) and len(want) == 36
    with pytest.raises(KeyError):
        paos_b200.pipeline({"conf": conf})  # save defaults to True, as in the reference: it needs passvalue['output']
    with pytest.raises(NotImplementedError):
        paos_b200.pipeline({"conf": conf, "save": False, "plot": True})


def test_wfe_column_ranges():
    from paos_b200.pipeline import parse_wfe_columns

    assert parse_wfe_columns("a.csv,7") == ("a.csv", [7])
    assert parse_wfe_columns("a.csv,3-6") == ("a.csv", [3, 4, 5, 6])
    assert parse_wfe_columns("dir/a.csv,0:10:4") == ("dir/a.csv", [0, 4, 8])
    assert parse_wfe_columns("a.csv,2.0") == ("a.csv", [2])


def test_bench_cpu_arm_keeps_the_full_affinity(monkeypatch):
    """bench.py narrows the GPU arm's process to the NUMA node of its GPU; the CPU workers of the baseline leg must count
    and use the CPUs the bench was started with (PAOS_BENCH_AFFINITY), not the narrowed set."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import importlib

    bench = importlib.import_module("bench")
    allowed = sorted(os.sched_getaffinity(0))
    monkeypatch.delenv("PAOS_BENCH_AFFINITY", raising=False)
    assert bench._full_affinity() == set(allowed) and bench.cpu_cores() == len(allowed)
    monkeypatch.setenv("PAOS_BENCH_AFFINITY", ",".join(str(c) for c in allowed[:1]))
    assert bench._full_affinity() == {allowed[0]} and bench.cpu_cores() == 1
    monkeypatch.setenv("PAOS_BENCH_AFFINITY", ",".join(str(c) for c in allowed))
    bench._restore_affinity()  # a no-op here; must not raise
    assert set(os.sched_getaffinity(0)) == set(allowed)
