"""CPU gates for the oracle:
  G0  oracle restatement == the reference TU compiled in place, bitwise (FP32 and FP64-typed)
  golden: oracle == committed vectors produced by running the reference (scripts/make_golden.py)
  KAT:  hand-derived / survey known answers
  portable-math provider (what the CUDA STRICT kernel reproduces bit-for-bit) is a rounding-level substitute
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, bit_equal, max_abs


def _params_from(po, z):
    over = {k: float(v) for k, v in z["params"]} if z["params"].size else {}
    return po.default_params(**over)


def _oracle_from(po, z, math=None, **kw):
    prec = int(z["prec"])
    p = _params_from(po, z)
    if prec == 64:
        # the FP64-typed reference widens its float-literal DEFAULTS; parameters written
        # afterwards through ref_set_param arrive as doubles, exactly like slider writes would
        over = {k: float(v) for k, v in z["params"]} if z["params"].size else {}
        p = po.float_rounded(po.default_params())
        for k, v in over.items():
            setattr(p, k, v)
    o = po.Oracle(int(z["nx"]), int(z["ny"]), p, prec=prec, math=po.MATH_LIBM if math is None else math, **kw)
    o.clear()
    for (x, y) in z["nuclei"]:
        o.add_nucleus(int(x), int(y))
    return o


GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_f*.npz")))      # field trajectories (ref_colors_* is the colour fixture)


def test_golden_files_present():
    assert len(GOLDEN_FILES) >= 9


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in GOLDEN_FILES])
def test_oracle_matches_reference_golden(po, path):
    """Oracle (libm provider, same compiler flags) reproduces the reference's fields bit-for-bit."""
    z = np.load(path, allow_pickle=True)
    o = _oracle_from(po, z)
    done = 0
    for s in z["steps"]:
        o.step(int(s) - done)
        done = int(s)
        phi, t, a = o.fields()
        assert bit_equal(phi, z[f"phi_{s}"]), f"phi differs at step {s}: {max_abs(phi, z[f'phi_{s}'])}"
        assert bit_equal(t, z[f"t_{s}"]), f"T differs at step {s}"
        assert bit_equal(a, z[f"angl_{s}"]), f"theta differs at step {s}"


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("j", [4.0, 6.0])
def test_g0_oracle_vs_reference_tu(po, prec, j):
    """G0: live run of the reference TU (only where oracle/_ref exists) vs the restatement, 250^2, 200 steps."""
    if not po.ref_available(prec):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = po.Reference(250, 250, float(np.float32(1e-4)) if prec == 64 else 1e-4, prec=prec, anisotropy=j)
    p = po.default_params(anisotropy=j)
    o = po.Oracle(250, 250, po.float_rounded(p) if prec == 64 else p, prec=prec)
    for n in (1, 9, 190):
        r.step(n)
        o.step(n)
        for a, b in zip(r.fields(), o.fields()):
            assert bit_equal(a, b)


def test_g0_iupdate_is_ten_substeps(po):
    if not po.ref_available(32):
        pytest.skip("oracle/_ref not built")
    r = po.Reference(64, 64, 1e-4)
    r.update()                      # Kobayashi::iUpdate, src/Kobayashi.cpp:227-239
    o = po.Oracle(64, 64)
    o.step(10)
    assert all(bit_equal(a, b) for a, b in zip(r.fields(), o.fields()))


def test_kat_first_substep_n32(po):
    """SURVEY §8c known answers, FP32, n=32, defaults, after sub-step 1 (c = 16)."""
    o = po.Oracle(32, 32)
    o.step(1)
    phi, t, th = o.fields()
    c = 16
    f = np.float32
    assert phi[c, c] == f(0.945555568) and t[c, c] == f(-0.0871110931)
    assert phi[c, c + 1] == f(0.868888915) and phi[c, c - 1] == f(0.868888915) and t[c, c + 1] == f(-0.209777743)
    assert phi[c + 1, c] == f(0.940493822) and phi[c - 1, c] == f(0.940493822) and t[c + 1, c] == f(-0.0952098891)
    assert phi[c + 1, c + 1] == f(0.0543209948) and t[c + 1, c + 1] == f(0.0869135931) and th[c + 1, c + 1] == f(3.92699075)
    assert phi[c, c + 2] == f(0.027222218) and phi[c + 2, c] == f(0.0148765482) and phi[c + 1, c + 2] == f(0.00249999762)
    assert phi[c, c + 3] == 0.0
    assert th[c, c + 1] == f(3.1415925) and th[c + 1, c] == f(-1.57079625) and th[c - 1, c] == f(1.57079625)
    # hand check: centre = 1 + (0.0105^2 * (-4/0.0027)) * (1e-4/3e-4)
    assert abs(float(phi[c, c]) - (1 + (0.0105 ** 2 * (-4 / 0.0027)) * (1e-4 / 3e-4))) < 1e-6
    assert abs(phi.astype(np.float64).sum() - 4.97469143) < 1e-8
    assert abs(t.astype(np.float64).sum() - (-0.0404937183)) < 1e-8
    o.step(1)
    phi, t, _ = o.fields()
    assert phi[c, c] == f(0.907500029) and abs(phi.astype(np.float64).sum() - 5.08597006) < 1e-8
    o.step(1)
    phi, t, _ = o.fields()
    assert phi[c, c] == f(0.880384564) and abs(t.astype(np.float64).sum() - 0.38377503) < 1e-8


def test_kat_sums_n250_long(po):
    """sums after 10/100/500/2000 sub-steps at 250^2 (default j=6), solid-cell counts (G6)."""
    o = po.Oracle(250, 250, threads=po.lib().kobo_max_threads())
    want = {10: (6.84774482, 2.95639169, 5), 100: (109.564119, 167.302589, 111), 500: (1350.16808, 2152.2692, 1334),
            2000: (10841.1094, 17337.7827, 10817)}
    done = 0
    for s, (sp, st, solid) in want.items():
        o.step(s - done)
        done = s
        phi, t, _ = o.fields()
        assert abs(phi.astype(np.float64).sum() / sp - 1) < 1e-8
        assert abs(t.astype(np.float64).sum() / st - 1) < 1e-8
        assert int((phi > 0.5).sum()) == solid


def test_openmp_threads_do_not_change_bits(po):
    a = po.Oracle(96, 80, threads=1)
    b = po.Oracle(96, 80, threads=4)
    a.step(40)
    b.step(40)
    assert all(bit_equal(x, y) for x, y in zip(a.fields(), b.fields()))


def test_extensions_are_bit_neutral_when_off(po):
    """theta0 = 0 and a = 0 leave the reference arithmetic untouched; a != 0 and theta0 != 0 change it."""
    base = po.Oracle(48, 48)
    base.step(30)
    off = po.Oracle(48, 48, po.default_params(theta0=0.0, noise_a=0.0), seed=99)
    off.step(30)
    assert all(bit_equal(x, y) for x, y in zip(base.fields(), off.fields()))
    on = po.Oracle(48, 48, po.default_params(noise_a=0.01), seed=99)
    on.step(30)
    assert not bit_equal(base.fields()[0], on.fields()[0])
    rot = po.Oracle(48, 48, po.default_params(theta0=0.3))
    rot.step(30)
    assert not bit_equal(base.fields()[0], rot.fields()[0])


def test_injected_noise_field_equals_philox_stream(po):
    """North star: 'an identical noise field injected from the host' — replay Philox draws as a host field."""
    nx, ny, seed = 40, 24, 1234
    p = po.default_params(noise_a=0.02)
    a = po.Oracle(nx, ny, p, seed=seed)
    b = po.Oracle(nx, ny, p, seed=0)
    for s in range(5):
        r = np.array([[po.noise_r(seed, s, i, j) for i in range(nx)] for j in range(ny)], np.float32)
        b.set_noise_field(r)
        a.step(1)
        b.step(1)
    assert all(bit_equal(x, y) for x, y in zip(a.fields(), b.fields()))


def test_portable_math_accuracy(po):
    """kob_math.h p_atan/p_sin/p_cos vs libm: a rounding-level substitute (few ulp)."""
    L = po.lib()
    xs = np.concatenate([np.linspace(-50, 50, 20001), np.linspace(-1e-3, 1e-3, 2001), [0.0, 1.0, -1.0, 2.4142135, 0.41421357]])
    for x in xs:
        x32 = float(np.float32(x))
        assert abs(L.kobo_p_atanf(x32) - np.arctan(x32)) <= 3e-7
        assert abs(L.kobo_p_sinf(x32) - np.sin(x32)) <= 3e-7
        assert abs(L.kobo_p_cosf(x32) - np.cos(x32)) <= 3e-7
        assert abs(L.kobo_p_atan(float(x)) - np.arctan(x)) <= 5e-16
        assert abs(L.kobo_p_sin(float(x)) - np.sin(x)) <= 5e-16
        assert abs(L.kobo_p_cos(float(x)) - np.cos(x)) <= 5e-16


def test_portable_oracle_tracks_reference_on_parity_windows(po):
    """The portable-math oracle (bit-equal to the CUDA STRICT kernel) vs reference vectors on the windows of
    SURVEY §4: FP32 single step from a warm state <= 1e-6, warm +100/+200 <= 1e-4; FP64 warm +500 <= 1e-10."""
    z = np.load(os.path.join(GOLDEN, "ref_f32_n128_j6_warm.npz"), allow_pickle=True)
    o = po.Oracle(128, 128, math=po.MATH_PORTABLE)
    o.set_fields(z["phi_500"], z["t_500"], z["angl_500"])
    o.step(1)
    assert max_abs(o.fields()[0], z["phi_501"]) <= 1e-6 and max_abs(o.fields()[1], z["t_501"]) <= 1e-6
    o.step(99)
    assert max_abs(o.fields()[0], z["phi_600"]) <= 1e-4
    o.step(100)
    assert max_abs(o.fields()[0], z["phi_700"]) <= 1e-4 and max_abs(o.fields()[1], z["t_700"]) <= 1e-4
    z = np.load(os.path.join(GOLDEN, "ref_f64_n128_j6_warm.npz"), allow_pickle=True)
    o = po.Oracle(128, 128, po.float_rounded(po.default_params()), prec=64, math=po.MATH_PORTABLE)
    o.set_fields(z["phi_500"], z["t_500"], z["angl_500"])
    o.step(1)
    assert max_abs(o.fields()[0], z["phi_501"]) <= 1e-14
    o.step(499)
    assert max_abs(o.fields()[0], z["phi_1000"]) <= 1e-10 and max_abs(o.fields()[1], z["t_1000"]) <= 1e-10
    # cold start, FP64-typed, 300 steps
    z = np.load(os.path.join(GOLDEN, "ref_f64_n64_j6.npz"), allow_pickle=True)
    o = po.Oracle(64, 64, po.float_rounded(po.default_params()), prec=64, math=po.MATH_PORTABLE)
    o.step(300)
    assert max_abs(o.fields()[0], z["phi_300"]) <= 1e-10


def test_philox_known_answers(po):
    """Random123 Philox4x32-10 KAT."""
    assert po.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert po.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert po.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_python_restatement_and_noise_mapping(po):
    """Independent pure-Python Philox + the (i>>2, j, step) -> word (i&3) -> (w>>8)*2^-24 mapping of kob_math.h."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def ph(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
            k = [(k[0] + W0) & 0xffffffff, (k[1] + W1) & 0xffffffff]
        return c

    seed, step = 0x1234567887654321, (7 << 32) | 5
    for (i, j) in [(0, 0), (1, 0), (2, 3), (3, 3), (4, 3), (16383, 16383), (65535, 70000)]:
        w = ph([i >> 2, j, step & 0xffffffff, step >> 32], [seed & 0xffffffff, seed >> 32])[i & 3]
        assert po.noise_r(seed, step, i, j) == (w >> 8) * 2.0 ** -24
    rs = np.array([po.noise_r(1, 0, i, j) for i in range(64) for j in range(64)])
    assert 0.0 <= rs.min() and rs.max() < 1.0 and abs(rs.mean() - 0.5) < 0.02


def test_nucleus_wraps_and_reset(po):
    o = po.Oracle(16, 12)
    o.clear()
    o.add_nucleus(0, 0)
    phi = o.fields()[0]
    assert phi[0, 0] == 1 and phi[0, 1] == 1 and phi[0, 15] == 1 and phi[1, 0] == 1 and phi[11, 0] == 1 and phi.sum() == 5
    o.reset()    # _vectorInit: nucleus at (nx/2, ny/2)
    phi = o.fields()[0]
    assert phi[6, 8] == 1 and phi[6, 7] == 1 and phi[6, 9] == 1 and phi[5, 8] == 1 and phi[7, 8] == 1 and phi.sum() == 5


def test_far_field_stays_exactly_zero(po):
    o = po.Oracle(64, 64)
    o.step(20)
    phi, t, th = o.fields()
    assert (phi[:8] == 0).all() and (t[:4] == 0).all() and (th[:8] == 0).all()


def test_colour_fixture_is_the_reference_ramp(po):
    """tests/golden/ref_colors_n48.npz (reference output, scripts/make_golden.py) against the ramp as the product restates
    it (src/Kobayashi.cpp:318-342: three linear segments with breaks at 0.9 and 0.99, transposed object -> cell mapping)
    — and, where oracle/_ref exists, against the reference TU run live on the same field."""
    z = np.load(os.path.join(GOLDEN, "ref_colors_n48.npz"))
    n, phi, rgb = int(z["n"]), z["phi"], z["rgb"]
    c0, c1, c2, c3 = (np.array(c, np.float32) for c in ([0, 0, 0], [0.2505490, 0.5, 0.9882353], [0.3607843, 1.0, 0.9882353], [0.9005490, 1.0, 0.9882353]))
    i = np.arange(n * n)
    p_ = phi[i % n, i // n]                                        # object i reads _phi[_INDEX(i / n, i % n)]
    want = np.empty((n * n, 3), np.float32)
    one = np.float32(1.0)
    for (lo, hi, a, b, sel) in ((0.0, 0.9, c0, c1, p_ <= np.float32(0.9)), (0.9, 0.99, c1, c2, (p_ > np.float32(0.9)) & (p_ <= np.float32(0.99))),
                                (0.99, 1.0, c2, c3, p_ > np.float32(0.99))):
        r = ((p_ - np.float32(lo)) * (one / (np.float32(hi) - np.float32(lo))))[:, None]
        want[sel] = (a * (one - r) + b * r)[sel]
    assert np.abs(want.astype(np.float64) - rgb.astype(np.float64)).max() <= 2e-6
    if po.ref_available(32):
        r = po.Reference(n, n, 1e-4, prec=32)
        r.set_fields(phi, np.zeros_like(phi), np.zeros_like(phi))
        assert np.array_equal(r.colors(), rgb)
