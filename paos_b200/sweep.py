"""Batch front-end: run many independent propagations on one GPU, or shard them over the GPUs of a box.

Replaces the reference's ``joblib.Parallel`` fan-out over wavelengths (``paos/core/pipeline.py:140-150``).
Jobs (see ``paos_b200/configs.py``) are independent, so the only parallel structure is:

* inside a GPU: a few *slots*, each a persistent :class:`WFO` on its own CUDA stream.  The host plans job k+1
  (pure scalar work) while the device still executes job k; the image-plane read-out of every job lands in a
  device stack, and (``host=True``) is copied to pinned host memory on the same stream, overlapping the next job;
* across GPUs: one process per GPU (``torch.distributed``), contiguous cost-balanced blocks of the job list per
  rank, no communication during propagation, and ONE gather of the result stack at the end
  (:func:`gather_stack`, NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os

import numpy as np

from . import _lib
from .run import run
from .wfo import WFO

READS = {"amplitude": _lib.READ_AMPLITUDE, "psf": _lib.READ_PSF, "phase": _lib.READ_PHASE}


def job_cost(job):
    """Relative cost of a job: number of surfaces that propagate (each costs 1-2 FFT2)."""
    return sum(1 for it in job["opt_chain"].values()
               if np.isfinite(it["ABCDt"].thickness) and abs(it["ABCDt"].thickness) > 1e-10) or 1


def partition(jobs, world_size):
    """Contiguous, cost-balanced split of ``jobs`` into ``world_size`` blocks -> list of (start, stop)."""
    n = len(jobs)
    cost = np.array([job_cost(j) for j in jobs], dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world_size):
        target = cum[-1] * r / world_size
        k = int(np.searchsorted(cum, target, side="left"))
        k = min(max(k, bounds[-1]), n)
        bounds.append(k)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def bind_to_gpu_numa(device):
    """Pin the calling process (and therefore the pinned host buffers it allocates next) to the CPUs of the NUMA
    node the GPU hangs off, so that device-to-host copies do not cross the socket interconnect.  Best effort:
    returns the NUMA node or None when the topology cannot be read."""
    import os

    try:
        import torch

        prop = torch.cuda.get_device_properties(device)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def default_batch(gridsize):
    """Wavefronts per batched launch.  A pass of one 2048^2 wavefront whose aperture blanks most lines is 270-820 CTAs, a
    wave or two of the 592 lines a B200 holds: several wavefronts per grid fill the machine and amortise the launch.
    Smaller grids need more items for the same effect; 4096^2 passes are already many waves."""
    n = int(gridsize)
    return 4 if n >= 4096 else 8 if n >= 2048 else 16


class Sweep:
    """Persistent executor for jobs of one grid size on one GPU.

    ``slots`` host threads / CUDA streams, each owning ``batch`` wavefronts: a slot takes ``batch`` consecutive jobs, plans
    their chains (``paos_batch_chain_run``) and the library runs them in lockstep, one kernel launch per pass for the whole
    batch.  ``batch=1`` is the unbatched executor (one wavefront per slot)."""

    def __init__(self, gridsize, device=0, dtype="complex128", slots=None, what="psf", batch=None):
        import torch

        if what not in READS:
            raise ValueError(f"what must be one of {sorted(READS)}")
        self.n = int(gridsize)
        cap = int(_lib.lib.paos_batch_capacity())
        if batch is None:
            batch = default_batch(self.n)
        self.batch = max(1, min(int(batch), cap))
        if slots is None:
            # batched: two slots are enough to overlap one batch's host planning and device-to-host copies with the
            # other's kernels (a third one covers the tails); unbatched (batch=1): see the round-1 measurements --
            # 2048^2 gains nothing beyond 4 slots, launch-bound small grids gain up to 6
            slots = 3 if self.batch > 1 else (4 if self.n >= 2048 else 6)
        self.device = int(device)
        self.dtype = dtype
        self.what = what
        self.torch = torch
        self.tdev = torch.device("cuda", self.device)
        self.rdtype = torch.float64 if dtype == "complex128" else torch.float32
        # PAOS_SWEEP_PRIORITIES=1 (off by default): slot s gets a higher stream priority than slot s + 1, so that the slots'
        # batches finish one after the other instead of all at once and the device-to-host copies of an end-to-end sweep
        # start after one batch time instead of `slots` of them.  Measured on one box (profiles/README.md): with
        # PAOS_SWEEP_FIRST_GROUP=2 as well the end-to-end sweep gains 6.5 % (1 415 -> 1 508 PSF/s), the device-resident one
        # loses 1.6 % (2 120 -> 2 087): the default stays with the device-resident rate.

        least, greatest = torch.cuda.Stream.priority_range()  # e.g. (0, -5): smaller is more urgent
        levels = abs(greatest - least)
        use_prio = os.environ.get("PAOS_SWEEP_PRIORITIES", "0") == "1" and slots > 1 and levels > 0
        step = -1 if greatest < least else 1
        self.streams = [torch.cuda.Stream(device=self.tdev, priority=(least + step * min(levels, slots - 1 - s)) if use_prio else least)
                        for s in range(slots)]
        # device-to-host copies of finished results run on a stream of their own, so that a slot's next batch of kernels
        # does not queue behind 8 x 32 MiB of PCIe traffic
        self.copy_streams = [torch.cuda.Stream(device=self.tdev) for _ in range(slots)]
        # wfos[s][i]: wavefront i of slot s; all wavefronts of a slot share the slot's stream
        self.slot_wfos = [[WFO(1.0, 1.0e-6, self.n, 1.0, device=self.device, dtype=dtype, stream=s) for _ in range(self.batch)]
                          for s in self.streams]
        self.wfos = [w for ws in self.slot_wfos for w in ws]
        self._pool = None
        self._screens = {}  # device copies of grid-sag maps shared between jobs

    @property
    def ring_rows(self):
        """Rows a result ring must hold a multiple of (every wavefront in flight owns one row)."""
        return len(self.streams) * self.batch

    def empty_stack(self, count, host=False):
        torch = self.torch
        if host:
            return torch.empty((count, self.n, self.n), dtype=self.rdtype, pin_memory=True)
        return torch.empty((count, self.n, self.n), dtype=self.rdtype, device=self.tdev)

    def run(self, jobs, out=None, host_out=None, psd_noise=None, native=True, cache_compiled=True, threads=True,
            ee=None, ee_out=None, ee_host_out=None, host_window=None, peak_out=None, on_group=None):
        """Propagate ``jobs``; the last saved surface of job k is read out (``what``) into ``out[k]``.

        ``ee``: ``dict(r_max=..., nbins=...)`` also reduces every PSF (``what="psf"``) to its encircled-energy curve on the
        device (``paos_b200/ee.py``) into ``ee_out[k]`` (``[len(jobs), nbins + 1]`` float64, allocated when None) and,
        if given, the pinned ``ee_host_out``; ``out`` may then be a ring of a multiple of ``ring_rows`` wavefronts, so that
        a sweep moves kilobytes per PSF instead of ``8 N^2`` bytes.  The curves are returned as ``meta[k]["ee"]`` views.

        ``out``: device stack (allocated when None).  ``host_out``: optional pinned host stack that also receives
        every result (asynchronous device-to-host copies inside the pipeline; a stack shorter than ``jobs`` is used
        as a ring of a multiple of ``ring_rows`` rows -- a row is valid once ``run`` has returned or been overwritten,
        which only suits a consumer that reads after the sweep).  ``native``: run each chain through ``paos_chain_run`` /
        ``paos_batch_chain_run`` (C++ per-surface loop) instead of the Python driver; ``cache_compiled=False`` rebuilds
        the native surface records of every job on every call.
        ``host_window``: ``(x0, y0, nx, ny)`` -- ``host_out`` then receives only that window of every read-out, narrowed to
        ``host_out``'s dtype (``paos_crop_convert``: a centred 512^2 float32 window of a 2048^2 PSF is 1 MiB on the PCIe link
        instead of 32 MiB).  ``peak_out``: ``[len(jobs), 2]`` float64 device tensor receiving the on-axis value and the
        maximum of every read-out (``paos_psf_peak``, the numerator of a Strehl ratio).  ``on_group(first_job, stop_job,
        event)``: called from the slot's thread after each batch has been enqueued, with a CUDA event recorded behind it
        (used to overlap the gather of finished PSFs with the rest of the sweep).
        Returns ``(out, meta)`` with one ``meta`` dict of host scalars per job, after all device work has completed.
        """
        import ctypes as C

        torch = self.torch
        if out is None:
            out = self.empty_stack(len(jobs))
        code = READS[self.what]
        nslots, B = len(self.streams), self.batch
        if out.shape[0] < len(jobs) and (ee is None or out.shape[0] % self.ring_rows):
            raise ValueError("out is shorter than jobs: only an encircled-energy sweep may use it as a ring, of a multiple of "
                             "`ring_rows` (= slots * batch) rows")
        if host_out is not None and host_out.shape[0] < len(jobs) and host_out.shape[0] % self.ring_rows:
            # rows are written by the wavefront that owns job k: a ring whose length is not a multiple of the wavefronts
            # in flight would let two streams copy into the same pinned row with no ordering between them
            raise ValueError("host_out is shorter than jobs: as a ring it must hold a multiple of `ring_rows` rows")
        if ee is not None:
            if self.what != "psf":
                raise ValueError("encircled energy needs what='psf'")
            from . import ee as ee_mod

            if ee_out is None:
                ee_out = torch.empty((len(jobs), int(ee["nbins"]) + 1), dtype=torch.float64, device=self.tdev)
        from . import chain as chain_mod

        meta = [None] * len(jobs)
        stage = None
        if host_window is not None:
            if host_out is None:
                raise ValueError("host_window needs host_out")
            x0, y0, nx, ny = (int(v) for v in host_window)
            if tuple(host_out.shape[1:]) != (ny, nx) or host_out.dtype not in (torch.float32, torch.float64):
                raise ValueError(f"host_out must be [rows, {ny}, {nx}] float32 or float64")
            key = (nx, ny, host_out.dtype)
            if getattr(self, "_stage_key", None) != key:
                self._stage = [[torch.empty((ny, nx), dtype=host_out.dtype, device=self.tdev) for _ in range(B)] for _ in range(nslots)]
                self._stage_key = key
            stage = self._stage

        def compile_native(k, job):
            """Compiled chain of job k with its read-out wired to out[k], or None when the native runner cannot take it."""
            if not native:
                return None
            if not cache_compiled:
                job.pop("_compiled", None)
            try:
                cc = chain_mod.compile_job(job, psd_noise(job) if psd_noise is not None else None, device=self.device,
                                           screen_cache=self._screens)
            except NotImplementedError:
                return None  # e.g. orthonormal Zernikes, resampled grid sag: the Python driver handles them
            if not cc.saved:
                raise ValueError(f"job {job.get('tag', k)} saves no surface")
            for idx in cc.saved:
                cc.set_readout(idx, -1, None)
            # the sweep keeps nothing but this read-out: the last pass need not store the complex field
            cc.set_readout(cc.saved[-1], code, out[k % out.shape[0]].data_ptr(), final=True)
            return cc

        def python_driver(k, job, wfo):
            dst = out[k % out.shape[0]]

            def snapshot(w, item):
                _lib.check(_lib.lib.paos_wfo_read_device(w._handle, code, C.c_void_p(dst.data_ptr())))
                return dict(wz=w.wz, distancetofocus=w.distancetofocus, fratio=w.fratio, dx=w.dx, dy=w.dy, wl=w.wl,
                            extent=w.extent, propagator=w.propagator)

            res = run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"],
                      job["opt_chain"], wfo=wfo, snapshot=snapshot, psd_noise=psd_noise(job) if psd_noise is not None else None)
            last = res[max(res.keys())] if res else {}
            return {k2: v for k2, v in last.items() if k2 not in ("aperture", "wfe")}

        # copies of whole read-outs into a host stack that holds every job go to the slot's copy stream (the source row of
        # `out` is written once per sweep); ring buffers and staged windows stay ordered on the slot's own stream
        side_copies = host_out is not None and stage is None and host_out.shape[0] >= len(jobs) and out.shape[0] >= len(jobs)
        pending = [[] for _ in range(nslots)]

        def finish(k, job, wfo, stream, last, s=0, i=0):
            dst = out[k % out.shape[0]]
            if peak_out is not None:
                _lib.check(_lib.lib.paos_psf_peak(wfo._handle, C.c_void_p(dst.data_ptr()), C.c_void_p(peak_out[k].data_ptr())))
            last["tag"] = job.get("tag", str(k))
            meta[k] = last
            if ee is not None:
                ee_mod.encircled_energy(wfo, dst, last["dx"], last["dy"], last["fratio"], last["wl"], ee.get("r_max", 8.0),
                                        ee["nbins"], out=ee_out[k])
                last["ee"] = ee_out[k]
                if ee_host_out is not None:
                    with torch.cuda.stream(stream):
                        ee_host_out[k].copy_(ee_out[k], non_blocking=True)
            if host_out is not None:
                if stage is not None:
                    st = stage[s][i]
                    _lib.check(_lib.lib.paos_crop_convert(wfo._handle, C.c_void_p(dst.data_ptr()), x0, y0, nx, ny,
                                                          1 if host_out.dtype == torch.float32 else 0, C.c_void_p(st.data_ptr())))
                    dst = st
                if side_copies:
                    pending[s].append((k, dst))
                else:
                    with torch.cuda.stream(stream):
                        host_out[k % host_out.shape[0]].copy_(dst, non_blocking=True)

        def do_group(s, ks):
            """Jobs ``ks`` (at most ``batch`` of them) on slot s."""
            wfos, stream = self.slot_wfos[s], self.streams[s]
            ccs = [compile_native(k, jobs[k]) for k in ks]
            nat = [i for i, cc in enumerate(ccs) if cc is not None]
            lasts = [None] * len(ks)
            if nat:
                try:
                    if len(nat) == 1:
                        i = nat[0]
                        lasts[i] = chain_mod.run_compiled(wfos[i], jobs[ks[i]], ccs[i])[-1]
                    else:
                        res = chain_mod.run_compiled_batch([wfos[i] for i in nat], [jobs[ks[i]] for i in nat], [ccs[i] for i in nat])
                        for i, r in zip(nat, res):
                            lasts[i] = r[-1]
                except NotImplementedError:
                    # a chain met a surface the native runner cannot take as compiled (a grid-sag map behind a change of
                    # sampling: its screen was resampled for the INIT pitch); the Python driver restarts these jobs from INIT
                    for i in nat:
                        lasts[i] = None
            for i, k in enumerate(ks):
                if lasts[i] is None:
                    lasts[i] = python_driver(k, jobs[k], wfos[i])
                finish(k, jobs[k], wfos[i], stream, lasts[i], s, i)
            if pending[s]:
                done = torch.cuda.Event()
                done.record(stream)
                cs = self.copy_streams[s]
                cs.wait_event(done)
                with torch.cuda.stream(cs):
                    for k, dst in pending[s]:
                        host_out[k].copy_(dst, non_blocking=True)
                pending[s].clear()
            if on_group is not None:
                ev = torch.cuda.Event()
                ev.record(stream)
                on_group(ks[0], ks[-1] + 1, ev)

        # the first group of every slot may be smaller (PAOS_SWEEP_FIRST_GROUP, experiment): results start to flow earlier
        first = int(os.environ.get("PAOS_SWEEP_FIRST_GROUP", "0") or 0)
        if out.shape[0] < len(jobs) or (host_out is not None and host_out.shape[0] < len(jobs)):
            first = 0  # rings rely on every slot owning the same rows group after group: equal groups only
        groups, g = [], 0
        while g < len(jobs):
            size = first if (0 < first < B and len(groups) < nslots) else B
            groups.append(list(range(g, min(g + size, len(jobs)))))
            g += size

        def do_slot(s):
            # one host thread per slot: the C++ planner and the launches run without the GIL (ctypes releases it),
            # so the slots' host work overlaps; the groups of a slot stay in order on the slot's stream
            torch.cuda.set_device(self.device)
            for g in range(s, len(groups), nslots):
                do_group(s, groups[g])

        distinct = len({id(j) for j in jobs}) == len(jobs)  # a job dict caches per-call native state
        if threads and distinct and nslots > 1 and len(groups) >= 2 * nslots:
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor

                self._pool = ThreadPoolExecutor(max_workers=nslots)
            for f in [self._pool.submit(do_slot, s) for s in range(nslots)]:
                f.result()
        else:
            for g, ks in enumerate(groups):
                do_group(g % nslots, ks)
        for stream in self.streams + self.copy_streams:
            stream.synchronize()
        return out, meta

    def stats(self):
        tot = {}
        for w in self.wfos:
            for k, v in w.stats().items():
                tot[k] = tot.get(k, 0) + v
        return tot

    def enable_timing(self, on=True):
        for w in self.wfos:
            _lib.check(_lib.lib.paos_wfo_enable_timing(w._handle, 1 if on else 0))

    def timing_detail(self, reset=True):
        """{(col, nfft): (ms, launches)} summed over the wavefronts (nfft: average chain length of a batched launch)."""
        import ctypes as C

        out = {}
        for w in self.wfos:
            _lib.check(_lib.lib.paos_wfo_sync(w._handle))
            for col in (0, 1):
                for nfft in range(0, _lib.MAX_CHAINED_FFTS + 1):
                    ms, cnt = C.c_double(), C.c_uint64()
                    _lib.check(_lib.lib.paos_wfo_timing_detail(w._handle, col, nfft, C.byref(ms), C.byref(cnt), 1 if reset else 0))
                    if cnt.value:
                        a = out.get((col, nfft), (0.0, 0))
                        out[(col, nfft)] = (a[0] + ms.value, a[1] + cnt.value)
        return out

    def timing_totals(self, reset=True):
        """Device time of the timed pass launches, their number, the line-FFT sweeps and the wavefront-passes they carried."""
        import ctypes as C

        tot = dict(ms=0.0, launches=0, line_fft_sweeps=0, wavefront_passes=0)
        for w in self.wfos:
            _lib.check(_lib.lib.paos_wfo_sync(w._handle))
            ms, a, b, c = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            _lib.check(_lib.lib.paos_wfo_timing_totals(w._handle, C.byref(ms), C.byref(a), C.byref(b), C.byref(c), 1 if reset else 0))
            tot["ms"] += ms.value
            tot["launches"] += a.value
            tot["line_fft_sweeps"] += b.value
            tot["wavefront_passes"] += c.value
        return tot


def gather_stack(local, counts, dst=0, group=None):
    """Gather per-rank stacks ``[n_local, N, N]`` on rank ``dst`` (the single collective of a sweep).

    ``counts``: jobs per rank (ragged blocks are padded to the largest count for the collective).  Returns the
    concatenated stack on ``dst`` and ``None`` elsewhere.  Works with NCCL (CUDA tensors) and gloo (CPU tensors).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    width = max(counts)
    pad = local
    if local.shape[0] != width:
        pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    pad = pad.contiguous()
    if rank == dst:
        # receive straight into one [world * width, N, N] stack (views as the gather list): no second copy when the
        # blocks are equal, which is the common case
        full = torch.empty((world * width,) + tuple(pad.shape[1:]), dtype=pad.dtype, device=pad.device)
        bufs = [full[r * width:(r + 1) * width] for r in range(world)]
        dist.gather(pad, gather_list=bufs, dst=dst, group=group)
        if all(c == width for c in counts):
            return full
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    dist.gather(pad, gather_list=None, dst=dst, group=group)
    return None


def chunk_schedule(counts, chunk, row_bytes):
    """The chunked gather as every rank must issue it: one entry per chunk g (jobs [g*chunk, (g+1)*chunk) of every rank's
    block) with the bytes each rank contributes and the byte offset of its rows in the destination stack, which holds the
    ranks' blocks back to back in rank order.  Ranks with fewer jobs contribute zero bytes to the last chunks (ragged
    blocks), so that the sequence of collective calls is the same everywhere."""
    counts = [int(c) for c in counts]
    chunk = max(1, int(chunk))
    base = [sum(counts[:q]) for q in range(len(counts))]
    out = []
    for g in range((max(counts) + chunk - 1) // chunk if counts else 0):
        lo = g * chunk
        sizes = [max(0, min(c, lo + chunk) - lo) * row_bytes for c in counts]
        offs = [(b + lo) * row_bytes for b in base]
        out.append((lo, sizes, offs))
    return out


class ChunkGather:
    """The sweep's one collective, overlapped with the sweep: every finished batch of PSFs is sent to its final rows of the
    destination stack on rank ``dst`` (``paos_gather_psf``: grouped NCCL send / receive over NVLink on a side stream) while
    later wavelengths still propagate.  Usage on every rank::

        cg = ChunkGather(device)                      # once: builds the communicator (collective)
        cg.begin(stack, counts, chunk=sweep.batch)    # per sweep: stack = local [n_local, N, N] device tensor
        sweep.run(jobs, out=stack, on_group=cg.on_group)
        full = cg.finish()                            # [sum(counts), N, N] on rank dst, None elsewhere

    NCCL calls of one communicator must be issued in the same order everywhere, so a single pump thread per rank sends the
    chunks in index order, whatever order the slots finish them in.
    """

    def __init__(self, device, dst=0, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        self.torch, self.C = torch, C
        self.rank, self.world, self.dst = dist.get_rank(group), dist.get_world_size(group), int(dst)
        self.device = int(device)
        ident = [None]
        if self.rank == 0:
            buf = (C.c_char * 128)()
            _lib.check_comm(_lib.lib.paos_comm_unique_id(C.cast(buf, C.c_void_p)))
            ident[0] = bytes(buf.raw)
        dist.broadcast_object_list(ident, src=0, group=group)
        h = C.c_void_p()
        idbuf = C.create_string_buffer(ident[0], 128)
        torch.cuda.set_device(self.device)
        _lib.check_comm(_lib.lib.paos_comm_create(C.byref(h), C.cast(idbuf, C.c_void_p), self.rank, self.world, self.device))
        self._comm = h
        self.stream = torch.cuda.Stream(device=torch.device("cuda", self.device))
        self._full = None

    def stack(self, counts, row_shape, dtype):
        """Local result stack ``[counts[rank], *row_shape]``.  On the destination rank it is a view into the full
        ``[sum(counts), *row_shape]`` stack at the rank's own offset, so its own block never has to be copied."""
        torch = self.torch
        dev = torch.device("cuda", self.device)
        counts = [int(c) for c in counts]
        if self.rank != self.dst:
            return torch.empty((counts[self.rank],) + tuple(row_shape), dtype=dtype, device=dev)
        self._full = torch.empty((sum(counts),) + tuple(row_shape), dtype=dtype, device=dev)
        lo = sum(counts[: self.rank])
        return self._full[lo: lo + counts[self.rank]]

    def close(self):
        h, self._comm = self._comm, None
        if h is not None:
            _lib.lib.paos_comm_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def begin(self, local, counts, chunk):
        import threading

        torch = self.torch
        self.local, self.counts, self.chunk = local, [int(c) for c in counts], max(1, int(chunk))
        self.row_bytes = local[0].numel() * local.element_size() if local.shape[0] else 0
        total = sum(self.counts)
        if self.rank == self.dst:
            if self._full is None or tuple(self._full.shape) != (total,) + tuple(local.shape[1:]) or self._full.dtype != local.dtype:
                self._full = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        self.schedule = chunk_schedule(self.counts, self.chunk, self.row_bytes)
        self._events = {}
        self._cv = threading.Condition()
        self._error = None
        self._thread = threading.Thread(target=self._pump, daemon=True)
        self._thread.start()

    def on_group(self, lo, hi, event):
        """Callback for ``Sweep.run(on_group=...)``: jobs [lo, hi) of the local block are done once ``event`` has fired."""
        with self._cv:
            self._events[lo // self.chunk] = event
            self._cv.notify_all()

    def _pump(self):
        C, torch = self.C, self.torch
        try:
            torch.cuda.set_device(self.device)
            mine = self.counts[self.rank]
            for g, (lo, nbytes, offsets) in enumerate(self.schedule):
                if lo < mine:
                    with self._cv:
                        while g not in self._events:
                            self._cv.wait()
                        ev = self._events[g]
                    self.stream.wait_event(ev)
                sizes = (C.c_size_t * self.world)()
                offs = (C.c_size_t * self.world)()
                for q in range(self.world):
                    sizes[q], offs[q] = nbytes[q], offsets[q]
                src = self.local[lo].data_ptr() if lo < mine else None
                dstp = self._full.data_ptr() if self.rank == self.dst else None
                _lib.check_comm(_lib.lib.paos_gather_psf(self._comm, C.c_void_p(src), sizes, offs, C.c_void_p(dstp), self.dst,
                                                         C.c_void_p(self.stream.cuda_stream)))
        except Exception as exc:  # surfaced by finish()
            self._error = exc

    def finish(self):
        self._thread.join()
        if self._error is not None:
            raise self._error
        self.stream.synchronize()
        return self._full if self.rank == self.dst else None
