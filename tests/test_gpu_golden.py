"""GPU path against the committed golden vectors (outputs of the unmodified reference, tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10  # BASELINE.json north_star: 1e-10 relative for complex128


def test_primitives_against_reference_golden():
    import paos_b200

    def setf(w, a):
        w.wfo = a

    def psd(w, noise, **kw):
        return w.psd(units="nm", noise=noise, **kw)

    golden = np.load(os.path.join(GOLDEN, "primitives.npz"))
    got = gc.primitives(lambda *a: paos_b200.WFO(*a), setf, lambda w: w.wfo, psd)
    assert set(got) == set(golden.files)
    gc.compare_to_golden(got, golden, "", TOL)


@pytest.mark.parametrize("case", range(6))
def test_chains_against_reference_golden(case, tmp_path):
    import paos_b200
    from paos_b200 import configs

    golden = np.load(os.path.join(GOLDEN, "chains.npz"))
    name, job, seed = gc.chain_jobs(str(tmp_path))[case]
    noise = configs.psd_noise_from_seed(seed) if seed is not None else None
    res = paos_b200.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"],
                        psd_noise=noise)
    got = gc.chain_summary(res)
    assert {f"{name}/{k}" for k in got} == {f for f in golden.files if f.startswith(name + "/")}
    gc.compare_to_golden(got, golden, name + "/", TOL)
