#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (the headline config is bench.py's): PSFs/s through
paos_b200.sweep.Sweep with results left in HBM, plus the wall time of ONE job on the CPU oracle for scale.

    python tools/config_bench.py [--cpu] [--only hubble,fgs1,ta_psd,grid_sag,airs512,airs1024,airs4096]
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="also time one job of each config on the numpy oracle")
    ap.add_argument("--only", default="")
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--slots", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="wavefronts per batched launch (default by grid size; 1 = unbatched)")
    args = ap.parse_args()
    import torch

    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    tmp = tempfile.mkdtemp(prefix="paos_cfg_")
    cases = {
        "hubble": lambda: ([configs.hubble(light_output=True)[0] for _ in range(64)], None),
        "airs512": lambda: (configs.airs_ch0(grid=512, n_wl=256), None),
        "airs1024": lambda: (configs.airs_ch0(grid=1024, n_wl=128), None),
        "airs4096": lambda: (configs.airs_ch0(grid=4096, n_wl=16), None),
        "fgs1": lambda: (configs.fgs1_montecarlo(grid=512, realizations=range(256)), None),
        "fgs1_2048": lambda: (configs.fgs1_montecarlo(grid=2048, realizations=range(32)), None),
        "ta_psd": lambda: (configs.ta_ground_psd(grid=1024, n_wl=8), "device_rng"),
        "grid_sag": lambda: (configs.grid_sag(grid=4096, wavelengths=(0.55, 3.0, 7.8) * 4, workdir=tmp), None),
    }
    only = [s for s in args.only.split(",") if s]
    out = {}
    for name, make in cases.items():
        if only and name not in only:
            continue
        jobs, mode = make()
        n = jobs[0]["gridsize"]
        sw = Sweep(n, slots=args.slots, what="psf", dtype=args.dtype, batch=args.batch)
        stack = sw.empty_stack(len(jobs))
        sw.run(jobs, out=stack)  # warm-up (also compiles the native surface records of every job: parse-time work)
        st0 = sw.stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sw.run(jobs, out=stack)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st1 = sw.stats()
        rec = {"grid": n, "jobs": len(jobs), "batch": sw.batch, "slots": len(sw.streams), "psf_per_s": len(jobs) / dt,
               "ms_per_psf": 1e3 * dt / len(jobs),
               "fft2_per_psf": (st1["fft2_recorded"] - st0["fft2_recorded"]) / len(jobs),
               "passes_per_psf": (st1["passes_planned"] - st0["passes_planned"]) / len(jobs),
               "pass_launches_per_psf": (st1["pass_launches"] - st0["pass_launches"]) / len(jobs),
               "launches_per_psf": (st1["kernel_launches"] - st0["kernel_launches"]) / len(jobs),
               "host_plan_us_per_psf": (st1["host_plan_us"] - st0["host_plan_us"]) / len(jobs)}
        rec["algorithmic_GBps"] = rec["fft2_per_psf"] * 64 * n * n * rec["psf_per_s"] / 1e9 * (1 if args.dtype == "complex128" else 0.5)
        if args.cpu:
            from oracle import paos_np

            job = jobs[0]
            noise = configs.psd_noise_from_seed(job["psd_seed"]) if "psd_seed" in job else None
            t0 = time.perf_counter()
            paos_np.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"],
                        noise_for=noise, unit_to_m=lambda u: u.to(type(u)("m")))
            rec["cpu_oracle_s_per_psf_1core"] = time.perf_counter() - t0
        # device side alone: CUDA events around every pass launch of one more sweep (they serialise the launches, so this
        # sweep is not the one timed above)
        sw.enable_timing(True)
        sw.timing_totals(reset=True)
        sw.run(jobs, out=stack)
        tot = sw.timing_totals(reset=True)
        sw.enable_timing(False)
        # launch durations summed over the slots' streams: launches of different slots overlap on the device, so this is the
        # device work per PSF, not a lower bound of the wall time
        rec["pass_kernel_us_per_psf"] = 1e3 * tot["ms"] / len(jobs)
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
        del sw, stack
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
