#!/usr/bin/env python
"""Benchmark of the PAOS Fresnel-propagation hot path on B200 (contract: see the task statement / DESIGN.md).

``python bench.py --gpus N --steps K --warmup W``           our arm (hand-written sm_100a kernels)
``python bench.py --impl reference --gpus N --steps K ...`` the reference's CPU algorithm on the host cores

Workload (BASELINE.json configs[1]): ``Ariel_AIRS-CH0.ini`` at 2048^2 complex128, 256 wavelengths 1.95-3.9 um,
only IMAGE_PLANE |.|^2 materialised.  A *step* is one full 256-wavelength sweep per GPU (weak scaling: rank r
sweeps field point r), metric = PSFs/s over all ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "2048^2 fp64 PSFs/sec (full surface chain)"
UNIT = "PSF/s"
GRID = 2048
N_WL = 256
# False: every timed sweep rebuilds the native surface records of every job from its opt_chain dictionary (what a first
# sweep over fresh jobs pays); --cache-compiled keeps them between sweeps (diagnostic: isolates the device side)
CACHE_COMPILED = False


def build_jobs(world, grid=GRID, n_wl=N_WL):
    """256 wavelengths x `world` field points (field 0 on-axis, field r at y = 0.01*r deg)."""
    import numpy as np
    from paos_b200 import configs

    base = configs.airs_ch0(grid=grid, n_wl=n_wl)
    jobs = []
    for r in range(world):
        ut = float(np.tan(np.deg2rad(0.01 * r)))
        for j in base:
            jj = dict(j)
            jj["field"] = {"us": j["field"]["us"], "ut": j["field"]["ut"] + ut}
            jj["tag"] = f"field{r}/" + j["tag"]
            jobs.append(jj)
    return jobs


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores.
#
# Preferred: the UNMODIFIED reference modules (``/root/reference`` in the build container, the byte-identical copy staged by
# ``tools/stage_ref.py`` under the git-ignored ``oracle/_ref/`` on the GPU box) through the reference's own public API --
# ``parse_config`` on a lens file, then ``run`` -- loaded by ``oracle/refload.py``: ``kind = "reference"``.  Fallback when
# no copy is present: the numpy port ``oracle/paos_np.py`` (bit-identical to the reference on the same numpy): "port".
# Neither the parent nor the workers import ``paos_b200`` (whose __init__ would dlopen the CUDA library).
# ---------------------------------------------------------------------------------------------------
LENS_FILE = os.path.join(ROOT, "paos_b200", "lens_data", "Ariel_AIRS-CH0.ini")


def sweep_wavelengths(n_wl):
    """The sweep of BASELINE.json configs[1] in micron (same expression as paos_b200.configs.airs_ch0)."""
    import numpy as np

    return np.linspace(1.95, 3.9, n_wl) if n_wl > 1 else np.array([1.95])


def cpu_kind():
    from oracle import refload

    return "reference" if refload.reference_available() else "port"


def _host_only_configs():
    """paos_b200.configs without running paos_b200/__init__.py (pure host-side parsing; no dlopen of the CUDA library)."""
    import importlib
    import types

    if "paos_b200" not in sys.modules:
        pkg = types.ModuleType("paos_b200")
        pkg.__path__ = [os.path.join(ROOT, "paos_b200")]
        sys.modules["paos_b200"] = pkg
    return importlib.import_module("paos_b200.configs")


def _cpu_worker(args):
    """One chain of the sweep on one core.  Returns (seconds, path of the saved IMAGE_PLANE amplitude or None)."""
    grid, n_wl, index, ut, keep_dir = args
    os.environ["OMP_NUM_THREADS"] = "1"
    _restore_affinity()
    import numpy as np

    from oracle import refload

    wl_um = float(sweep_wavelengths(n_wl)[index])
    field = {"us": 0.0, "ut": float(ut)}
    if refload.reference_available():
        import configparser
        import tempfile

        ref = refload.load()
        cfg = configparser.ConfigParser()
        cfg.read(LENS_FILE)
        cfg["general"]["grid_size"] = str(grid)
        for key in list(cfg["wavelengths"]):
            del cfg["wavelengths"][key]
        cfg["wavelengths"]["w1"] = repr(wl_um)
        with tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False) as fh:
            cfg.write(fh)
            path = fh.name
        try:
            pup, params, wls, fields, chains = ref.parse_config(path)
        finally:
            os.unlink(path)
        chain = chains[0]
        for item in chain.values():  # pipeline.py:112-115 light_output: keep the IMAGE_PLANE snapshot only
            item["save"] = item["name"] == "IMAGE_PLANE"
        f0 = dict(fields[0])
        f0["ut"] = f0["ut"] + field["ut"]
        t0 = time.perf_counter()
        res = ref.run(pup, 1.0e-6 * wls[0], params["grid_size"], params["zoom"], f0, chain)
    else:
        from oracle import paos_np

        job = _host_only_configs().airs_ch0(grid=grid, n_wl=n_wl)[index]
        f0 = {"us": job["field"]["us"], "ut": job["field"]["ut"] + field["ut"]}
        t0 = time.perf_counter()
        res = paos_np.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], f0, job["opt_chain"])
    dt = time.perf_counter() - t0
    out = None
    if keep_dir:
        out = os.path.join(keep_dir, f"amp_{index}.npy")
        np.save(out, res[max(res.keys())]["amplitude"])
    return dt, out


def _full_affinity():
    """The CPUs this run may use: the affinity the bench was started with (the GPU arm narrows its own to one NUMA node)."""
    spec = os.environ.get("PAOS_BENCH_AFFINITY")
    if spec:
        return {int(c) for c in spec.split(",") if c}
    try:
        return set(os.sched_getaffinity(0))
    except AttributeError:
        return set(range(os.cpu_count() or 1))


def _restore_affinity():
    try:
        os.sched_setaffinity(0, _full_affinity())
    except (AttributeError, OSError):
        pass


def cpu_cores():
    return len(_full_affinity()) or 1


def cpu_chains(pool, indices, grid=GRID, n_wl=N_WL, ut=0.0, keep_dir=None):
    """Propagate the wavelengths ``indices`` of the sweep over the pool.  Returns (PSF/s, seconds, worker results)."""
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(grid, n_wl, int(i), ut, keep_dir) for i in indices], chunksize=1)
    dt = time.perf_counter() - t0
    return len(indices) / dt, dt, res


def spread(count, n_wl, offset=0):
    """``count`` wavelength indices spread evenly over the sweep."""
    return [(offset + i * max(1, n_wl // count)) % n_wl for i in range(count)]


def make_pool(procs):
    import multiprocessing as mp

    return mp.get_context("spawn").Pool(procs)


def cpu_procs():
    # one process per core, capped: every worker holds ~1.5 GB of 2048^2 complex128 temporaries
    return max(1, min(cpu_cores(), int(os.environ.get("PAOS_BENCH_CPU_PROCS", "32"))))


def run_reference(args):
    """The CPU arm as a stand-alone run.  A 2048^2 chain is ~30-50 s of numpy on one core and cannot be made shorter, so
    the run is sized from a measured chain time: one round (one chain per core) is timed first; it is the warm-up when
    warm-up steps were asked for.  The timed region is then ONE ``pool.map`` of R rounds (R * procs chains, every core busy
    throughout) with R chosen so that the whole run stays inside ``PAOS_BENCH_REF_BUDGET_S`` (default 420 s); it is
    reported as ``steps`` equal steps of R * procs / steps chains each.  PSF/s = chains / wall time does not depend on
    how the chains are cut into steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t_start = time.perf_counter()
    procs = cpu_procs()
    budget_s = float(os.environ.get("PAOS_BENCH_REF_BUDGET_S", "420"))
    kind = cpu_kind()
    pool = make_pool(procs)
    try:
        _, t_round, _ = cpu_chains(pool, spread(procs, N_WL, 0))
        warm_done = 1
        used = time.perf_counter() - t_start
        rounds = int(max(1, min(args.steps, (budget_s - used) // (1.1 * t_round))))
        idx = []
        for r in range(rounds):
            idx += spread(procs, N_WL, 1 + r)
        value, dt, _ = cpu_chains(pool, idx)
    finally:
        pool.close()
        pool.join()
    n_chains = len(idx)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Ariel_AIRS-CH0.ini {GRID}^2 complex128, {N_WL} wavelengths 1.95-3.9 um per GPU "
                               f"(rank r = field point r), IMAGE_PLANE |.|^2 only",
                   "grid": GRID, "wavelengths_per_gpu": N_WL,
                   "sample": f"{n_chains} wavelengths spread evenly over that sweep, propagated by {procs} host processes in one "
                             f"pool.map ({rounds} rounds of one chain per core), reported as {args.steps} equal steps; "
                             "PSF/s = wavelengths / wall time"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{n_chains} chains in {dt:.1f} s on {procs} cores after {warm_done} untimed round "
                                   f"({t_round:.1f} s); budget {budget_s:.0f} s, whole run {time.perf_counter() - t_start:.0f} s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for ln in self.proc.stdout:
            if self.t_mark is not None and time.perf_counter() >= self.t_mark:
                self.lines.append(ln.strip())

    def mark(self):
        """Start of the timed region: nvidia-smi has been running since start() (it needs a few 100 ms to come up), only
        samples taken from now on are kept."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def timed_steps(torch, sweep, jobs, stack, host_stack, steps, extra_streams=(), before=None, after=None, **run_kw):
    """Device time (ms) of `steps` sweeps: events on the current stream fenced against every slot stream (and
    `extra_streams`: the gather's side stream); `before()` / `after()` bracket every sweep inside the timed region."""
    cur = torch.cuda.current_stream()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record(cur)
    for s in list(sweep.streams) + list(getattr(sweep, "copy_streams", [])) + list(extra_streams):
        s.wait_event(start)
    for _ in range(steps):
        if before is not None:
            before()
        sweep.run(jobs, out=stack, host_out=host_stack, cache_compiled=CACHE_COMPILED, **run_kw)
        if after is not None:
            after()
    for s in list(sweep.streams) + list(getattr(sweep, "copy_streams", [])) + list(extra_streams):
        e = torch.cuda.Event()
        e.record(s)
        cur.wait_event(e)
    end.record(cur)
    torch.cuda.synchronize()
    return start.elapsed_time(end)


# FP64 instructions of one 2048-point complex128 line FFT in the 16 x 8 x 16 form of fft_core.cuh, counted in the SASS of
# the pass kernel: per thread 320 DADD + 138 DFMA + 74 DMUL = 532 (butterflies and twiddles only: the per-position table
# multiply, the twiddle products and the mask factors are real work but are NOT counted), 128 threads per line.
FP64_INSTR_PER_LINE_FFT = {(2048, "complex128"): 532 * 128}


def load_peaks():
    """Measured ceilings: MEASURED_PEAKS.json (driver-written: HBM copy GB/s, max SM clock) and profiles/peaks_r02.json
    (tools/peaks.cu on this pool's B200: FP64 issue rate, L2 and shared-memory bandwidth)."""
    out = {}
    for name in ("MEASURED_PEAKS.json", os.path.join("profiles", "peaks_r02.json")):
        try:
            with open(os.path.join(ROOT, name)) as fh:
                out.update(json.load(fh))
        except (OSError, ValueError):
            pass
    return out


def fp64_peak(peaks):
    """(thread-instructions/s, source)."""
    rates = [float(peaks[k]) for k in ("dadd_Ginstr_s", "dmul_Ginstr_s", "dfma_Ginstr_s") if k in peaks]
    if rates:
        # the butterflies are 60 % DADD, 26 % DFMA, 14 % DMUL; the highest of the three measured issue rates is the ceiling
        return max(rates) * 1e9, "profiles/peaks_r02.json: highest measured FP64 issue rate (tools/peaks.cu: DADD/DMUL 63.8, DFMA 58.8 instr/clk/SM)"
    clk = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    return 64 * 148 * clk, "nominal 64 FP64 instr/clk/SM x 148 SMs x max SM clock (no measured probe found)"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import paos_b200
    from paos_b200 import sweep as sweep_mod

    grid, n_wl = args.grid, args.n_wl
    jobs_all = build_jobs(world, grid, n_wl)
    blocks = sweep_mod.partition(jobs_all, world)
    lo, hi = blocks[rank]
    jobs = jobs_all[lo:hi]
    counts = [b - a for a, b in blocks]

    # pin this process (and the pinned host buffers it allocates next) to the NUMA node of its GPU; the CPU workers of the
    # baseline leg go back to the full affinity recorded here (PAOS_BENCH_AFFINITY, applied in _cpu_worker)
    os.environ.setdefault("PAOS_BENCH_AFFINITY", ",".join(str(c) for c in sorted(os.sched_getaffinity(0))))
    numa = sweep_mod.bind_to_gpu_numa(local_rank) if (world > 1 or not os.environ.get("PAOS_BENCH_NO_NUMA")) else None
    sw = sweep_mod.Sweep(grid, device=local_rank, dtype=args.dtype, slots=args.slots, what="psf", batch=args.batch)
    args.slots, args.batch = len(sw.streams), sw.batch
    # multi-GPU: the local stack of the destination rank is a view into the gathered [world * n, N, N] stack
    cg = sweep_mod.ChunkGather(local_rank, dst=0) if world > 1 else None
    stack = cg.stack(counts, (grid, grid), sw.rdtype) if cg is not None else sw.empty_stack(len(jobs))
    host_ring = False
    try:
        host_stack = sw.empty_stack(len(jobs), host=True)
    except RuntimeError:
        # not enough pinned memory for the whole stack on this host: stream the PSFs through a 64-slot ring
        host_stack = sw.empty_stack(min(len(jobs), 2 * sw.ring_rows), host=True)
        host_ring = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") -------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        sw.run(jobs, out=stack)
    barrier()
    st0 = sw.stats()
    sampler.mark()
    t_host0 = time.perf_counter()
    ms = timed_steps(torch, sw, jobs, stack, None, args.steps)
    t_host = time.perf_counter() - t_host0
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    ms = max_over_ranks(ms)
    st1 = sw.stats()
    total_psf = len(jobs_all) * args.steps
    value = total_psf / (ms * 1e-3)

    # ---- the single collective: gather of the PSF stack to rank 0 (paos_gather_psf, NCCL over NVLink) -------------
    # (a) alone, after a sweep (gather_ms: what it costs when nothing hides it); (b) e2e_gathered: the sweep with the
    # gather of every finished batch overlapped on a side stream, both inside the timed region
    gather_ms = gather_first_ms = None
    e2e_gathered = None
    if world > 1:
        gathers = []
        for _ in range(2):  # the first call builds NCCL's point-to-point connections
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            g0.record(cg.stream)
            cg.begin(stack, counts, chunk=max(counts))
            ev = torch.cuda.Event()
            ev.record()
            cg.on_group(0, len(jobs), ev)
            cg.finish()
            g1.record(cg.stream)
            torch.cuda.synchronize()
            gathers.append(max_over_ranks(g0.elapsed_time(g1)))
        gather_first_ms, gather_ms = gathers
        barrier()
        ms_g = max_over_ranks(timed_steps(torch, sw, jobs, stack, None, args.steps, extra_streams=[cg.stream],
                                          before=lambda: cg.begin(stack, counts, chunk=sw.batch), after=cg.finish,
                                          on_group=cg.on_group))
        e2e_gathered = {"value": total_psf / (ms_g * 1e-3), "unit": UNIT + " (sweep + overlapped gather of every PSF to rank 0)",
                        "gathered_bytes_per_step": int(sum(counts[1:]) * grid * grid * (8 if args.dtype == "complex128" else 4)),
                        "gather_alone_ms": gather_ms}
        barrier()

    # ---- strong scaling (BASELINE.json configs[1] read literally: ONE 256-wavelength sweep sharded by wavelength) ----
    strong = None
    if world > 1:
        base = jobs_all[: n_wl]  # field point 0
        sblocks = sweep_mod.partition(base, world)
        slo, shi = sblocks[rank]
        sjobs = [dict(j) for j in base[slo:shi]]
        sw.run(sjobs, out=stack[: len(sjobs)])
        barrier()
        ms_s = max_over_ranks(timed_steps(torch, sw, sjobs, stack[: len(sjobs)], None, args.steps))
        strong = {"value": n_wl * args.steps / (ms_s * 1e-3), "unit": UNIT, "scaling": "strong",
                  "psf_per_step": n_wl, "psf_per_gpu_per_step": len(sjobs), "ms_per_step": ms_s / args.steps}
        barrier()

    # ---- end to end through the public API with host buffers ("e2e") -------------------------------
    sw.run(jobs, out=stack, host_out=host_stack)
    barrier()
    ms_e2e = max_over_ranks(timed_steps(torch, sw, jobs, stack, host_stack, args.steps))
    e2e_value = total_psf / (ms_e2e * 1e-3)
    d2h = int(len(jobs) * grid * grid * (8 if args.dtype == "complex128" else 4))
    barrier()

    # ---- the same sweep with a reduced host product: the centred grid/4 window of every PSF as float32 (1 MiB instead of
    # 32 MiB at 2048^2), cropped and narrowed on the device (paos_crop_convert) -----------------------------------------
    win = grid // 4
    small = torch.empty((len(jobs), win, win), dtype=torch.float32, pin_memory=True)
    window = ((grid - win) // 2, (grid - win) // 2, win, win)
    sw.run(jobs, out=stack, host_out=small, host_window=window)
    barrier()
    ms_red = max_over_ranks(timed_steps(torch, sw, jobs, stack, small, args.steps, host_window=window))
    e2e_reduced = {"value": total_psf / (ms_red * 1e-3), "unit": UNIT + f" (centred {win}^2 float32 window of every PSF to host)",
                   "d2h_bytes_per_step": int(len(jobs) * win * win * 4) * world}
    del small
    barrier()

    # ---- the same sweep when its product is the encircled-energy curve of every PSF: the PSFs stay on the
    # device in a ring of `slots` buffers and 2 KB per PSF crosses PCIe instead of 32 MiB --------------------------------
    e2e_ee = None
    if args.dtype == "complex128":
        nb = 256
        ring = stack[: sw.ring_rows]
        ee_dev = torch.empty((len(jobs), nb + 1), dtype=torch.float64, device=stack.device)
        ee_host = torch.empty((len(jobs), nb + 1), dtype=torch.float64, pin_memory=True)

        def ee_steps(steps):
            cur = torch.cuda.current_stream()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(cur)
            for s_ in sw.streams:
                s_.wait_event(a)
            for _ in range(steps):
                sw.run(jobs, out=ring, ee=dict(r_max=8.0, nbins=nb), ee_out=ee_dev, ee_host_out=ee_host, cache_compiled=CACHE_COMPILED)
            for s_ in sw.streams:
                e_ = torch.cuda.Event()
                e_.record(s_)
                cur.wait_event(e_)
            b.record(cur)
            torch.cuda.synchronize()
            return a.elapsed_time(b)

        ee_steps(1)
        barrier()
        ms_ee = max_over_ranks(ee_steps(args.steps))
        e2e_ee = {"value": total_psf / (ms_ee * 1e-3), "unit": "PSF/s (encircled-energy curves to host)",
                  "d2h_bytes_per_step": int(len(jobs) * (nb + 1) * 8) * world}
        barrier()

    # ---- per-kernel timing for the roofline (separate pass: events around every launch) -------------
    # The pass kernel is bound by the FP64 pipe (fused passes chain ~5 line FFTs per sweep of the
    # field, so the HBM model of SURVEY 8d no longer describes it; ncu: ~10 MB of DRAM traffic per launch).  achieved =
    # FP64 butterfly instructions of the lines really transformed / summed launch durations (CUDA events on the launching
    # stream, one slot so that launches do not overlap); peak = the measured DFMA issue rate of tools/peaks.cu.
    roofline = None
    passes = {}
    if rank == 0:
        sw1 = sweep_mod.Sweep(grid, device=local_rank, dtype=args.dtype, slots=1, what="psf", batch=args.batch)
        sw1.enable_timing(True)
        sub = jobs[: min(len(jobs), 8 * sw1.batch)]
        sw1.run(sub, out=stack[: len(sub)])
        sw1.timing_detail(reset=True)
        sw1.timing_totals(reset=True)
        s1a = sw1.stats()
        sw1.run(sub, out=stack[: len(sub)])
        det = sw1.timing_detail(reset=True)
        tot = sw1.timing_totals(reset=True)
        s1b = sw1.stats()
        sw1.enable_timing(False)
        del sw1
        elem = 16 if args.dtype == "complex128" else 8
        half_fft2_bytes = 2 * elem * grid * grid  # one read + one write of the field: half of SURVEY 8(d)'s 64 N^2
        tot_ms, tot_launch = tot["ms"], tot["launches"]
        alg = tot["line_fft_sweeps"] * half_fft2_bytes
        peaks = load_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get("pass_kernel_bytes_per_launch")
        except (OSError, ValueError):
            pass
        for (col, nfft), (m, c) in sorted(det.items()):
            passes[f"{'col' if col else 'row'}x{nfft}"] = {"launches": c, "avg_us": 1e3 * m / c}
        lines = int(s1b["lines_transformed"] - s1a["lines_transformed"])
        sweeps = int(s1b["line_ffts_run"] - s1a["line_ffts_run"])
        per_line = FP64_INSTR_PER_LINE_FFT.get((grid, args.dtype))
        peak_i, peak_src = fp64_peak(peaks)
        hbm_model = {"achieved_GBps": alg / (tot_ms * 1e-3) / 1e9 if tot_ms else None, "peak_GBps": hbm_peak,
                     "x_hbm_peak": alg / (tot_ms * 1e-3) / 1e9 / hbm_peak if tot_ms else None,
                     "algorithmic_bytes_per_launch": alg / max(tot_launch, 1),
                     "note": "SURVEY 8d convention: 32*N^2 B per line-FFT sweep (64*N^2 per FFT2) / launch time.  Not a "
                             "roofline fraction: fused passes chain several line FFTs per sweep of the field and blanked "
                             "lines are never touched, so this exceeds the HBM peak by construction"}
        # The pipe that binds next to FP64 (ncu, profiles/: l1tex__data_pipe_lsu_wavefronts 69 % on the batched row pass): the
        # L1 / shared-memory data path, 128 B/clk/SM.  Algorithmic bytes through it per line FFT: two exchanges, each element
        # written and read once (4 x sizeof(complex) x N); per tabled line N x sizeof(complex); per line FFT the twiddles
        # (11 loads of sizeof(complex) per thread, N/16 threads at 2048); per swept line one load and one store of the field.
        smem_view = None
        if tot_ms and "smem_GBs" in peaks:
            tabled = int(s1b["lines_tabled"] - s1a["lines_tabled"])
            swept = int(s1b["lines_swept"] - s1a["lines_swept"])
            bytes_l1 = lines * (4 * elem * grid + 11 * elem * (grid // 16)) + tabled * elem * grid + swept * 2 * elem * grid
            ach_l1 = bytes_l1 / (tot_ms * 1e-3) / 1e9
            smem_view = {"achieved_GBps": ach_l1, "peak_GBps": float(peaks["smem_GBs"]), "frac": ach_l1 / float(peaks["smem_GBs"]),
                         "bytes_per_launch": bytes_l1 / max(tot_launch, 1),
                         "note": "algorithmic bytes through the L1/shared-memory data pipe (exchanges, phase tables, twiddles, field) / "
                                 "launch time against the measured LDS+STS rate of tools/peaks.cu; ncu reads 69 % for the same pipe"}
        if per_line and tot_ms:
            ach = lines * per_line / (tot_ms * 1e-3)
            wall_ach = (int(st1["lines_transformed"] - st0["lines_transformed"]) * per_line) / (ms * 1e-3)
            roofline = {"bound": "fp64", "achieved": ach / 1e9, "peak": peak_i / 1e9, "unit": "Ginstr/s", "frac": ach / peak_i,
                        "traffic": traffic, "kernel": "pass_kernel (row and column instantiations, batched launches)",
                        "avg_launch_us": 1e3 * tot_ms / max(tot_launch, 1), "launches_timed": tot_launch,
                        "wavefronts_per_launch": tot["wavefront_passes"] / max(tot_launch, 1),
                        "fp64_instr_per_line_fft": per_line, "lines_transformed": lines,
                        "active_line_fraction": lines / max(sweeps * grid, 1),
                        "peak_source": peak_src,
                        "wall": {"achieved": wall_ach / 1e9, "frac": wall_ach / peak_i,
                                 "note": "same count over the whole timed region of `value` (all slots overlapping, every kernel included)"},
                        "l1_smem_pipe": smem_view, "algorithmic_model_x_hbm": hbm_model,
                        "note": "FP64 butterfly instructions (SASS count, table/mask multiplies excluded) of the lines really "
                                "transformed / pass-kernel launch time; the kernel keeps its working set in registers and "
                                "shared memory, DRAM traffic is `traffic` bytes per launch (ncu)"}
        else:
            roofline = {"bound": "hbm", "achieved": hbm_model["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_model["x_hbm_peak"], "traffic": traffic, "kernel": "pass_kernel",
                        "avg_launch_us": 1e3 * tot_ms / max(tot_launch, 1),
                        "note": hbm_model["note"] + "; no FP64 instruction count is tabulated for this grid / precision"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) + parity of the same wavelengths in the same run -------------
    # (BASELINE.md 4.4 / SURVEY 8d: max|amp_gpu - amp_ref| / max|amp_ref| <= 1e-10 for complex128, 1e-4 for complex64)
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import shutil
        import tempfile

        procs = cpu_procs()
        # the sample: the wavelengths whose chain takes the other propagator route (35 instead of 39 FFT2: different
        # inside/outside decisions of the pilot beam), then wavelengths spread evenly over the sweep, one per core
        sw1 = sweep_mod.Sweep(grid, device=local_rank, dtype=args.dtype, slots=1, what="amplitude", batch=1)
        nfft2 = []
        for j in jobs:
            f0 = sw1.stats()["fft2_recorded"]
            sw1.run([j], out=stack[:1])
            nfft2.append(int(sw1.stats()["fft2_recorded"] - f0))
        common = max(set(nfft2), key=nfft2.count)
        odd = [i for i, c in enumerate(nfft2) if c != common]
        sample = odd[:procs]
        for i in spread(procs, n_wl):
            if len(sample) < procs and i not in sample:
                sample.append(i)
        keep = tempfile.mkdtemp(prefix="paos_bench_cpu_")
        pool = make_pool(procs)
        try:
            v, dt, res = cpu_chains(pool, sample, grid, n_wl, keep_dir=keep)
            tol = 1e-10 if args.dtype == "complex128" else 1e-4
            amp_gpu, _ = sw1.run([jobs[i] for i in sample], out=stack[: len(sample)])
            worst, worst_psf, worst_at = 0.0, 0.0, None
            for k, (i, (_, path)) in enumerate(zip(sample, res)):
                ref_amp = torch.from_numpy(np.load(path)).to(stack.device)
                got = amp_gpu[k].to(torch.float64)
                err = float((got - ref_amp).abs().max() / ref_amp.abs().max())
                ref_psf = ref_amp * ref_amp
                err_psf = float((got * got - ref_psf).abs().max() / ref_psf.max())
                if err > worst:
                    worst, worst_at = err, i
                worst_psf = max(worst_psf, err_psf)
            parity = {"n": len(sample), "worst": worst, "worst_psf": worst_psf, "tolerance": tol, "ok": bool(worst <= tol),
                      "worst_wavelength_index": worst_at, "fft2_per_psf": {str(c): nfft2.count(c) for c in sorted(set(nfft2))},
                      "other_route_indices_checked": [i for i in odd if i in sample],
                      "what": "max|amp_gpu - amp_ref| / max|amp_ref| at IMAGE_PLANE over the cpu_baseline wavelengths "
                              "(worst_psf: the same for |.|^2), GPU vs the CPU arm's arrays of this run"}
        finally:
            pool.close()
            pool.join()
            shutil.rmtree(keep, ignore_errors=True)
        del sw1
        cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": cpu_kind(),
               "sample": f"{len(sample)} wavelengths of the same sweep (the {len([i for i in odd if i in sample])} with the other "
                         f"propagator route + evenly spread ones), one numpy process per core, {dt:.1f} s wall"}
        if not parity["ok"]:
            raise SystemExit(f"PARITY FAILURE inside bench.py: {json.dumps(parity)}")

    if rank == 0:
        launches = int(st1["kernel_launches"] - st0["kernel_launches"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.dtype == "complex128" else "f32", "data": "synthetic",
            "config": {"workload": f"Ariel_AIRS-CH0.ini {grid}^2 {args.dtype}, {n_wl} wavelengths 1.95-3.9 um per GPU "
                                   f"(rank r = field point r), IMAGE_PLANE |.|^2 only",
                       "grid": grid, "wavelengths_per_gpu": n_wl, "psf_per_step": len(jobs_all), "slots": args.slots,
                       "batch": args.batch,
                       "l2": f"each wavefront is {(16 if args.dtype == 'complex128' else 8) * grid * grid >> 20} MiB and "
                             f"{args.slots * args.batch} are in flight (> 126 MB L2 at 2048^2); every PSF is a different "
                             "wavelength, no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h * world,
                    "note": "paos_b200.sweep.Sweep.run over host job dicts; inputs are lens-prescription scalars (kernel "
                            "arguments, no array uploads); every PSF (N*N fp64) is copied to pinned host memory inside the timed region"},
            "gpu_launches": launches * world,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "numa_node": numa,
            "wall_ms_per_psf": 1e3 * t_host / (len(jobs) * args.steps),
            "host_plan_ms_per_psf": 1e-3 * (st1["host_plan_us"] - st0["host_plan_us"]) / (len(jobs) * args.steps),
            "fft2_per_step": int(st1["fft2_recorded"] - st0["fft2_recorded"]) // args.steps * world,
            "pass_launches_per_psf": (st1["pass_launches"] - st0["pass_launches"]) / (len(jobs) * args.steps),
            "passes_per_psf": (st1["passes_planned"] - st0["passes_planned"]) / (len(jobs) * args.steps),
            "e2e_ee": e2e_ee, "e2e_reduced": e2e_reduced, "e2e_gathered": e2e_gathered, "strong_scaling": strong,
            "passes": passes, "gather_ms": gather_ms, "gather_first_ms": gather_first_ms,
        }
        emit(line)
    if cg is not None:
        cg.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (--config): same contract, their own workload line
# ---------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (grid, what the workload line says, job builder(world) -> (jobs, psd noise provider or None))
    "hubble": (1024, "Hubble_simple.ini as shipped (1024^2, 1.0 um, on-axis), IMAGE_PLANE |.|^2 only, 256 propagations per GPU",
               lambda cfg, world: ((lambda base: [dict(base) for _ in range(256 * world)])(cfg.hubble(light_output=True)[0]), None)),
    "fgs1": (512, "Ariel_FGS-FGS1.ini + wfe_realization_SN20210914.csv, 256 Monte-Carlo WFE realizations per GPU at 512^2, 36 Zernike terms",
             lambda cfg, world: (cfg.fgs1_montecarlo(grid=512, realizations=range(256 * world)), None)),
    "ta_psd": (1024, "lens_file_TA_Ground_PSD.ini, 9 fields x 28 wavelengths per GPU at 1024^2, PSD screens drawn on the device (Philox)",
               lambda cfg, world: ([j for _ in range(world) for j in cfg.ta_ground_psd(grid=1024, n_wl=28)], None)),
    "grid_sag": (4096, "test_Grid_Sag.ini, synthetic on-grid sag at 4096^2 complex128, 12 wavelengths per GPU (map prepared on the device)",
                 lambda cfg, world: (cfg.grid_sag(grid=4096, wavelengths=tuple([0.55, 3.0, 7.8] * (4 * world))), None)),
}


def run_config(args):
    """One of the non-headline BASELINE.json configurations through the same front-end: device-resident PSF/s (`value`), end to
    end with every PSF copied to pinned host memory (`e2e`), parity of a few jobs against the CPU arm in the same run."""
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from paos_b200 import configs as cfg
    from paos_b200 import sweep as sweep_mod

    grid, workload, make = CONFIGS[args.config]
    jobs_all, noise = make(cfg, world)
    blocks = sweep_mod.partition(jobs_all, world)
    lo, hi = blocks[rank]
    jobs = jobs_all[lo:hi]
    sw = sweep_mod.Sweep(grid, device=local_rank, dtype=args.dtype, slots=args.slots, what="psf", batch=args.batch)
    stack = sw.empty_stack(len(jobs))
    host = sw.empty_stack(len(jobs), host=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        sw.run(jobs, out=stack)
    barrier()
    st0 = sw.stats()
    sampler.mark()
    ms = max_over_ranks(timed_steps(torch, sw, jobs, stack, None, args.steps))
    clocks = sampler.stop() if rank == 0 else None
    st1 = sw.stats()
    sw.run(jobs, out=stack, host_out=host)
    barrier()
    ms_e2e = max_over_ranks(timed_steps(torch, sw, jobs, stack, host, args.steps))
    total = len(jobs_all) * args.steps
    # the same sweep when the native surface records of the jobs are kept between sweeps (Sweep.run's default): at small grids
    # the Python record builder (~0.25 ms per job under the GIL) is what a sweep over fresh jobs waits for, not the device
    global CACHE_COMPILED
    keep, CACHE_COMPILED = CACHE_COMPILED, True
    sw.run(jobs, out=stack)
    barrier()
    ms_cached = max_over_ranks(timed_steps(torch, sw, jobs, stack, None, args.steps))
    CACHE_COMPILED = keep
    parity = None
    if rank == 0 and not args.no_cpu and args.config != "ta_psd":  # device-drawn PSD noise has no CPU counterpart (statistical mode)
        from oracle import paos_np

        pick = [0, len(jobs) // 2] if grid <= 1024 else [0]
        worst = 0.0
        for k in pick:
            j = jobs[k]
            ref = paos_np.run(j["pupil_diameter"], j["wavelength"], j["gridsize"], j["zoom"], j["field"], j["opt_chain"])
            psf_ref = ref[max(ref)]["amplitude"] ** 2
            worst = max(worst, float(np.max(np.abs(host[k].numpy() - psf_ref)) / np.max(psf_ref)))
        tol = 1e-10 if args.dtype == "complex128" else 1e-4
        parity = {"n": len(pick), "worst_psf": worst, "tolerance": tol, "ok": bool(worst <= tol), "against": "oracle/paos_np.py (port)"}
        if not parity["ok"]:
            raise SystemExit(f"PARITY FAILURE inside bench.py --config {args.config}: {json.dumps(parity)}")
    if rank == 0:
        rbytes = 8 if args.dtype == "complex128" else 4
        emit({
            "metric": f"{grid}^2 fp64 PSFs/sec (full surface chain, {args.config})", "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if args.dtype == "complex128" else "f32", "data": "synthetic",
            "config": {"workload": workload, "grid": grid, "psf_per_step": len(jobs_all), "slots": len(sw.streams), "batch": sw.batch},
            "clocks": clocks,
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(len(jobs_all) * grid * grid * rbytes)},
            "gpu_launches": int(st1["kernel_launches"] - st0["kernel_launches"]) * world,
            "passes_per_psf": (st1["passes_planned"] - st0["passes_planned"]) / (len(jobs) * args.steps),
            "pass_launches_per_psf": (st1["pass_launches"] - st0["pass_launches"]) / (len(jobs) * args.steps),
            "value_records_kept": {"value": total / (ms_cached * 1e-3), "unit": UNIT,
                                   "note": "compiled surface records kept between sweeps (Sweep.run default); `value` rebuilds them every sweep"},
            "parity": parity, "roofline": None, "cpu_baseline": None,
        })
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def claim_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.  Point fd 1
    at stderr for the duration of the run and keep the real stdout for emit()."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="complex128", choices=["complex128", "complex64"])
    ap.add_argument("--grid", type=int, default=GRID)
    ap.add_argument("--n-wl", type=int, default=N_WL)
    ap.add_argument("--slots", type=int, default=None, help="host threads / streams (default: 3 batched, 4 unbatched)")
    ap.add_argument("--batch", type=int, default=None, help="wavefronts per batched launch (default by grid size; 1 = unbatched)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--config", default="airs_ch0", choices=["airs_ch0"] + sorted(CONFIGS),
                    help="workload: airs_ch0 = BASELINE.json configs[1] (the headline, default); the others are its configs[0], [2], [3], [4]")
    ap.add_argument("--cache-compiled", action="store_true", help="diagnostic: keep the compiled surface records between sweeps")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch ourselves one rank per GPU (the driver normally does this with torchrun)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    global CACHE_COMPILED
    CACHE_COMPILED = bool(args.cache_compiled)
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "airs_ch0":
        return run_config(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
