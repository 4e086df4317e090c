"""Device-side noise mode (csrc/plant/rng.h, nps_set_device_rng).

CPU: the generator itself - Philox4x32-10 against the published known-answer vectors (through an independent Python
restatement of the round function) and the statistics of the five per-step draws, evaluated by the library's own host
entry point nps_device_rng_draws (no GPU needed).
GPU: a batch stepped with device-side noise equals the host oracle fed with the same generator's draws, and does not
depend on how the steps are grouped into launches or the plants into shards."""
import ctypes

import numpy as np
import pytest

from tests import _util as U


def _philox4x32_10(ctr, key):
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def test_python_restatement_matches_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert _philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    pi_case = _philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])
    assert pi_case == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _draws(lib, seed, plant, step):
    out = np.zeros(5)
    assert lib.nps_device_rng_draws(ctypes.c_uint64(seed), ctypes.c_uint64(plant), ctypes.c_uint64(step), U.ptr(out)) == 0
    return out


def _u53(a, b):
    return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0


def test_library_draws_follow_the_documented_construction():
    from nuclear_sim_b200 import _clib
    lib = _clib.lib()
    for seed, plant, step in [(0, 0, 0), (42, 7, 3), (2 ** 63 + 12345, 65535, 3599), (1, 2 ** 33 + 5, 2 ** 32 + 9)]:
        key = [seed & 0xFFFFFFFF, seed >> 32]
        base = [plant & 0xFFFFFFFF, plant >> 32, step & 0xFFFFFFFF, (step >> 32) & 0x3FFFFFFF]
        a = _philox4x32_10(base, key)
        b = _philox4x32_10(base[:3] + [base[3] | 0x40000000], key)
        e = _philox4x32_10(base[:3] + [base[3] | 0x80000000], key)
        u1, u2 = _u53(a[0], a[1]), _u53(a[2], a[3])
        r, phi = np.sqrt(-2.0 * np.log(1.0 - u1)), 2.0 * np.pi * u2
        want = np.array([r * np.cos(phi), r * np.sin(phi), _u53(b[0], b[1]), _u53(b[2], b[3]), _u53(e[0], e[1])])
        got = _draws(lib, seed, plant, step)
        np.testing.assert_allclose(got[:2], want[:2], rtol=1e-14, atol=1e-15)
        np.testing.assert_array_equal(got[2:], want[2:])


def test_draw_statistics():
    from nuclear_sim_b200 import _clib
    lib = _clib.lib()
    z = np.array([_draws(lib, 2024, p, s) for p in range(200) for s in range(100)])
    n = len(z)
    for j in (0, 1):                                  # standard normals
        assert abs(z[:, j].mean()) < 4 / np.sqrt(n) and abs(z[:, j].var() - 1) < 0.05
        assert abs((z[:, j] ** 4).mean() - 3) < 0.3
    for j in (2, 3, 4):                               # uniforms on [0, 1)
        assert z[:, j].min() >= 0 and z[:, j].max() < 1
        assert abs(z[:, j].mean() - 0.5) < 4 / np.sqrt(12 * n) and abs(z[:, j].var() - 1 / 12) < 0.005
    c = np.corrcoef(z.T)
    assert np.abs(c - np.eye(5)).max() < 0.03         # the five draws of a step are uncorrelated
    by_plant = z[:, 0].reshape(200, 100)
    assert abs(np.corrcoef(by_plant[0], by_plant[1])[0, 1]) < 0.3 and not np.array_equal(by_plant[0], by_plant[1])


@pytest.mark.gpu
def test_device_rng_steps_equal_oracle_with_the_same_draws(oracle_lib):
    import torch
    from nuclear_sim_b200 import BatchedNuclearPlantSimulator, _clib, load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    lib = _clib.lib()
    n, steps, seed, offset = 48, 12, 987654321, 1000
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")       # constant heat source with noise: z_heat matters
    st = sc.randomized_states(s0, np.arange(offset, offset + n))
    sim = BatchedNuclearPlantSimulator(n, st, params, device="cuda:0")
    sim.set_device_rng(seed, plant_offset=offset)
    sim.step(K=5); sim.step(K=4); sim.step(K=3)                  # 12 steps in uneven launches
    torch.cuda.synchronize()
    noise = np.array([[_draws(lib, seed, offset + p, s) for s in range(steps)] for p in range(n)])   # [n][steps][5]
    ref = np.ascontiguousarray(st, dtype=np.float64).copy()
    assert oracle_lib.nps_oracle_step(U.ptr(ref), U.ptr(np.ascontiguousarray(params)), None, None,
                                      U.ptr(np.ascontiguousarray(noise)), ctypes.c_int64(n), steps) == 0
    U.assert_states_close(sim.state_numpy(), ref, U.TOL_STEP * steps, "device rng vs oracle")
    # the noise really entered: a run without it differs
    quiet = BatchedNuclearPlantSimulator(n, st, params, device="cuda:0")
    quiet.step(K=steps)
    assert not np.array_equal(quiet.state_numpy(), sim.state_numpy())
    # sharding / launch grouping invariance: the second half of the plants alone, one launch
    half = BatchedNuclearPlantSimulator(n // 2, st[n // 2:], params, device="cuda:0")
    half.set_device_rng(seed, plant_offset=offset + n // 2)
    half.step(K=steps)
    np.testing.assert_array_equal(half.state_numpy(), sim.state_numpy()[n // 2:])
