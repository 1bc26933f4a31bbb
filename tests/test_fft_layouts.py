"""CPU model of the line FFT's exchange-buffer layouts (paos_b200/csrc/fft_core.cuh, in-place exchanges).

The kernel's claim: a three-stage transform N = E * R2 * E needs only the two barriers that follow its writes of the
exchange buffer, because every write goes to addresses the same thread has just read -- provided consecutive transforms
alternate between the two layouts tabulated in fft_core.cuh.  This test restates the six address maps in numpy, runs
chained transforms through a model of the buffer, and checks (1) the transform, (2) the in-place property the barrier
count rests on, (3) that every 16-byte (complex128) shared-memory phase of 8 consecutive threads hits 8 bank groups.
"""
import numpy as np
import pytest


def line_fft_model(v, n, e, flip, sm, log):
    """v[j, t] = x[t + j*T] -> X in the same distribution; records every (kind, addresses[register, thread])."""
    t_count = n // e
    r2 = n // (e * e)
    tp = t_count + 1
    nb = e // r2
    t = np.arange(t_count)
    n3, q = t % e, t // e
    k = np.arange(e)[:, None]
    v = np.fft.fft(v, axis=0) * np.exp(-2j * np.pi * k * t[None, :] / n)  # stage 1 + twiddle W_N^(k1 t)
    a0, a1 = t, n3 * tp + q * e
    w1 = (a0[None, :] + k * tp) if not flip else (a1[None, :] + k)
    log.append(("w", w1))
    sm[w1] = v
    c = np.arange(nb)[:, None, None]
    n2 = np.arange(r2)[None, :, None]
    mid = ((q * tp + n3)[None, None, :] + c * r2 * tp + n2 * e) if not flip else ((n3 * tp + q)[None, None, :] + c * r2 + n2 * e)
    log.append(("r", mid.reshape(e, t_count)))
    u = np.fft.fft(sm[mid], axis=1) * np.exp(-2j * np.pi * np.arange(r2)[None, :, None] * n3[None, None, :] / t_count)
    log.append(("w", mid.reshape(e, t_count)))
    sm[mid] = u
    last = (a1[None, :] + k) if not flip else (a0[None, :] + k * tp)
    log.append(("r", last))
    return np.fft.fft(sm[last], axis=0)


@pytest.mark.parametrize("n,e", [(512, 8), (1024, 16), (2048, 16), (4096, 16)])
def test_in_place_exchange_layouts(n, e):
    t_count = n // e
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    v = x.reshape(e, t_count).copy()
    sm = np.full(e * (t_count + 1) + 8, np.nan + 0j)
    log, ref = [], x
    for i in range(5):  # layouts alternate: flip = 0, 1, 0, 1, 0
        v = line_fft_model(v, n, e, i & 1, sm, log)
        ref = np.fft.fft(ref)
        assert np.allclose(v.reshape(-1), ref, rtol=0, atol=1e-9 * np.abs(ref).max()), (n, e, i)
    # every write after the very first goes to exactly the addresses the same thread read last: no barrier in front of it
    for (ka, a), (kb, b) in zip(log[1:], log[2:]):
        if ka == "r" and kb == "w":
            assert all(set(a[:, tt]) == set(b[:, tt]) for tt in range(t_count)), (n, e)
    # the barriers that remain separate a write from reads of OTHER threads: every read set differs from the writer's
    for (ka, a), (kb, b) in zip(log, log[1:]):
        if ka == "w" and kb == "r":
            assert any(set(a[:, tt]) != set(b[:, tt]) for tt in range(t_count))
    # complex128: a 16-byte access is served in phases of 8 consecutive threads; 8 distinct 16-byte bank groups each
    for _, a in log:
        for j in range(a.shape[0]):
            groups = a[j].reshape(-1, 8) % 8
            assert (np.sort(groups, axis=1) == np.arange(8)).all(), (n, e, j)
