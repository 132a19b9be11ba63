"""GPU parity (run on the B200 box, `-m gpu`): libkobayashi_cuda.so, called through its C ABI (ctypes mirror
class crystalgrowth_b200.Kobayashi), against the CPU oracle and the committed reference vectors.

  STRICT kernel  == oracle built with the portable math provider     : BITWISE, any number of steps (G3)
  STRICT kernel  vs reference golden vectors (libm)                   : tolerance windows of SURVEY §4 (G1/G2)
  FAST kernel    vs reference golden vectors / libm oracle            : FP32 1e-6 single step, 1e-4 on the windows
  full-size grids: size-independent properties (far field exactly 0, grid-size-independent sums, symmetry,
                   shard invariance, Philox replay)
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bit_equal, max_abs

pytestmark = pytest.mark.gpu

NUCLEI_EDGE = [(1, 1), (94, 38), (50, 0), (0, 20), (95, 39)]


def _pair(po, cg, nx, ny, prec="f32", kernel="strict", seed=7, nuclei=None, math=None, **params):
    p = po.default_params(**params)
    o = po.Oracle(nx, ny, p, prec=64 if prec == "f64" else 32, math=po.MATH_PORTABLE if math is None else math,
                  seed=seed, threads=po.lib().kobo_max_threads())
    g = cg.Kobayashi(nx, ny, 1e-4, precision=prec, kernel=kernel, seed=seed,
                     **{("t_eq" if k == "t_eq" else k): v for k, v in params.items()})
    if nuclei is not None:
        o.clear()
        g.clear()
        for (x, y) in nuclei:
            o.add_nucleus(x, y)
            g.add_nucleus(x, y)
    return o, g


def _assert_bitwise(o, g, what=""):
    for name, a, b in zip(("phi", "T", "theta"), g.fields(), o.fields()):
        assert bit_equal(a, b), f"{what}: {name} differs, max abs {max_abs(a, b)}"


# ---------------------------------------------------------------------------------------------- STRICT
@pytest.mark.parametrize("nx,ny,steps,prec,params,nuclei", [
    (64, 64, 1, "f32", {}, None),
    (64, 64, 60, "f32", {}, None),
    (250, 250, 400, "f32", {}, None),                                   # through the FP32-chaotic regime
    (250, 250, 400, "f32", {"anisotropy": 4.0}, None),
    (37, 53, 200, "f32", {"anisotropy": 5.0}, [(1, 1), (35, 51), (18, 26)]),   # ragged: no tile divides it
    (96, 40, 120, "f32", {"noise_a": 0.01}, NUCLEI_EDGE),                # nuclei on the periodic seams + Philox
    (96, 40, 80, "f32", {"theta0": 0.3, "anisotropy": 4.0}, NUCLEI_EDGE),
    (33, 2, 30, "f32", {}, [(5, 0), (20, 1)]),                           # thinnest legal grid in y
    (2, 31, 30, "f32", {}, [(0, 7)]),                                    # thinnest legal grid in x
    (300, 300, 120, "f32", {"K": 1.2, "tau": 0.0004, "delta": 0.03, "alpha": 1.1, "gamma": 15.0, "t_eq": 0.9,
                            "epsilon_bar": 0.012, "anisotropy": 8.0}, None),
    (64, 64, 150, "f64", {}, None),
    (130, 70, 150, "f64", {"noise_a": 0.02}, [(0, 0), (64, 35), (129, 69)]),
    (250, 250, 300, "f64", {"anisotropy": 4.0}, None),
    (1024, 1024, 20, "f32", {}, None),
])
def test_strict_is_bitwise_equal_to_oracle(po, cg, nx, ny, steps, prec, params, nuclei):
    o, g = _pair(po, cg, nx, ny, prec, "strict", nuclei=nuclei, **params)
    o.step(steps)
    g.step(steps)
    _assert_bitwise(o, g, f"{nx}x{ny} {prec} {steps} steps")


def test_strict_non_integer_anisotropy_and_step_by_step(po, cg):
    o, g = _pair(po, cg, 80, 72, "f32", "strict", anisotropy=5.5)
    for _ in range(25):
        o.step(1)
        g.step(1)
        _assert_bitwise(o, g)


def test_strict_injected_noise_field(po, cg):
    """North star: identical noise field injected from the host -> same result as the device Philox stream."""
    nx, ny, seed = 72, 40, 4242
    o, g = _pair(po, cg, nx, ny, "f32", "strict", seed=seed, noise_a=0.02)
    _, h = _pair(po, cg, nx, ny, "f32", "strict", seed=1, noise_a=0.02)
    for s in range(6):
        r = np.array([[po.noise_r(seed, s, i, j) for i in range(nx)] for j in range(ny)], np.float32)
        h.set_noise_field(r)
        h.step(1)
        g.step(1)
        o.step(1)
    _assert_bitwise(o, g)
    assert all(bit_equal(a, b) for a, b in zip(g.fields(), h.fields()))
    h.set_noise_field(None)


@pytest.mark.parametrize("kernel", ["strict", "fast"])
def test_warm_windows_vs_reference_vectors_f32(po, cg, kernel):
    """G1/G2 against vectors produced by the reference itself (glibc libm): warm start at step 500."""
    z = np.load(os.path.join(GOLDEN, "ref_f32_n128_j6_warm.npz"), allow_pickle=True)
    g = cg.Kobayashi(128, 128, 1e-4, precision="f32", kernel=kernel)
    g.set_fields(z["phi_500"], z["t_500"], z["angl_500"])
    g.step(1)
    phi, t, th = g.fields()
    assert max_abs(phi, z["phi_501"]) <= 1e-6 and max_abs(t, z["t_501"]) <= 1e-6     # G1
    g.step(99)
    assert max_abs(g.phi(), z["phi_600"]) <= 1e-4                                      # G2
    g.step(100)
    assert max_abs(g.phi(), z["phi_700"]) <= 1e-4 and max_abs(g.t(), z["t_700"]) <= 1e-4


@pytest.mark.parametrize("kernel", ["strict", "fast"])
@pytest.mark.parametrize("j", [4, 6])
def test_cold_window_vs_reference_vectors_f32(po, cg, kernel, j):
    """FP32 cold start: the nucleus-centre cell has a mathematically zero gradient, so its angle is decided by
    rounding noise against the FLT_EPSILON dead-band; with glibc libm that first flips phi by ~8e-3 at sub-step 7
    (SURVEY §5.7), with any other atan/sin/cos (portable provider, CUDA libm) it can flip at sub-step 6.  The
    tolerance window that every rounding-level variant satisfies is therefore N <= 4-5; N = 4 is tested."""
    z = np.load(os.path.join(GOLDEN, f"ref_f32_n64_j{j}.npz"), allow_pickle=True)
    g = cg.Kobayashi(64, 64, 1e-4, precision="f32", kernel=kernel, anisotropy=float(j))
    g.step(4)
    phi, t, _ = g.fields()
    assert max_abs(phi, z["phi_4"]) <= 1e-6 and max_abs(t, z["t_4"]) <= 1e-6


def test_first_substeps_vs_reference_vectors_all_fields(po, cg):
    z = np.load(os.path.join(GOLDEN, "ref_f32_n32_j6.npz"), allow_pickle=True)
    for kernel in ("strict", "fast"):
        g = cg.Kobayashi(32, 32, 1e-4, precision="f32", kernel=kernel)
        for s in (1, 2, 3):
            g.step(1)
            phi, t, th = g.fields()
            assert max_abs(phi, z[f"phi_{s}"]) <= 1e-6 and max_abs(t, z[f"t_{s}"]) <= 1e-6
            # theta is compared after the FIRST sub-step only (identical inputs -> identical decisions).  Later the
            # nucleus-centre cell (mathematically zero gradient) is assigned or held depending on 1-ulp noise in
            # phi, so theta there legitimately differs between rounding-level variants.
            if s == 1:
                d = np.abs(th.astype(np.float64) - z[f"angl_{s}"].astype(np.float64))
                d = np.minimum(d, np.abs(d - 2 * 3.1415926))
                assert d.max() <= 1e-5, kernel


def test_f64_vs_reference_vectors(po, cg):
    """FP64: 1e-10 after 300 cold steps and after +500 warm steps (BASELINE.json tolerance)."""
    z = np.load(os.path.join(GOLDEN, "ref_f64_n64_j6.npz"), allow_pickle=True)
    p = po.float_rounded(po.default_params())
    g = cg.Kobayashi(64, 64, p.dt, precision="f64", params=_as_cg_params(cg, p))
    g.step(300)
    assert max_abs(g.phi(), z["phi_300"]) <= 1e-10 and max_abs(g.t(), z["t_300"]) <= 1e-10
    z = np.load(os.path.join(GOLDEN, "ref_f64_n128_j6_warm.npz"), allow_pickle=True)
    g = cg.Kobayashi(128, 128, p.dt, precision="f64", params=_as_cg_params(cg, p))
    g.set_fields(z["phi_500"], z["t_500"], z["angl_500"])
    g.step(1)
    assert max_abs(g.phi(), z["phi_501"]) <= 1e-14
    g.step(499)
    assert max_abs(g.phi(), z["phi_1000"]) <= 1e-10 and max_abs(g.t(), z["t_1000"]) <= 1e-10


def _as_cg_params(cg, p):
    q = cg.KobParams()
    for name, _ in cg.KobParams._fields_:
        setattr(q, name, getattr(p, name))
    return q


# ---------------------------------------------------------------------------------------------- FAST
@pytest.mark.parametrize("nx,ny,params,nuclei", [
    (250, 250, {}, None),
    (250, 250, {"anisotropy": 4.0}, None),
    (96, 40, {"noise_a": 0.01}, NUCLEI_EDGE),
    (37, 53, {"anisotropy": 5.0}, [(1, 1), (35, 51), (18, 26)]),
    (33, 2, {}, [(5, 0), (20, 1)]),
    (2, 31, {}, [(0, 7)]),
    (200, 120, {"anisotropy": 3.0, "theta0": 0.2}, None),
    (200, 120, {"anisotropy": 5.5}, None),                               # non-integer mode: trig path
    (300, 300, {"K": 1.2, "tau": 0.0004, "delta": 0.03, "alpha": 1.1, "gamma": 15.0, "t_eq": 0.9,
                "epsilon_bar": 0.012, "anisotropy": 8.0}, None),
])
def test_fast_single_steps_from_oracle_states(po, cg, nx, ny, params, nuclei):
    """G1 for the roofline kernel: from identical (phi, T, theta) states along an oracle trajectory (cold AND
    evolved), one FAST step stays within FP32 1e-6 of one reference-arithmetic step."""
    o, g = _pair(po, cg, nx, ny, "f32", "fast", nuclei=nuclei, math=po.MATH_LIBM, **params)
    for advance in (0, 1, 4, 45, 150):
        o.step(advance)
        phi, t, th = o.fields()
        g.set_fields(phi, t, th)
        g.step_counter = o.step_counter()
        o.step(1)
        g.step(1)
        gp, gt, gth = g.fields()
        op, ot, oth = o.fields()
        assert max_abs(gp, op) <= 1e-6 and max_abs(gt, ot) <= 2e-6, f"after {o.step_counter()} steps"
        # theta: identical assignment decisions; values equal up to atan rounding, modulo 2*pi at the branch cut
        d = np.abs(gth.astype(np.float64) - oth.astype(np.float64))
        d = np.minimum(d, np.abs(d - 2 * 3.1415926))
        assert d.max() <= 2e-5


@pytest.mark.parametrize("j,theta0", [(6.0, 0.0), (4.0, 0.0), (5.0, 0.3), (5.5, 0.0)])
def test_fast_dense_field_steps(po, cg, j, theta0):
    """The bench's dense workload in small: every cell on a diffuse interface, so the packed data-dependent block
    (minimax atan for the angle and m(T), trig-free anisotropy, the shared Philox draw, held angles by MUFU) runs
    for every cell, in interior (live) and seam jobs alike.  Single FAST steps from oracle states stay within
    FP32 1e-6; the smooth field is not rounding-chaotic, so a 40-step trajectory holds 1e-4 as well."""
    import bench
    nx, ny = 420, 200
    phi, t = bench.dense_state(nx, ny, 0)
    p = po.default_params(noise_a=0.01, anisotropy=j, theta0=theta0)
    o = po.Oracle(nx, ny, p, prec=32, math=po.MATH_LIBM, seed=21)
    g = cg.Kobayashi(nx, ny, 1e-4, kernel="fast", seed=21, noise_a=0.01, anisotropy=j, theta0=theta0)
    z = np.zeros((ny, nx), np.float32)
    o.set_fields(phi, t, z)
    for advance in (0, 1, 7, 30, 80):
        o.step(advance)
        g.set_fields(*o.fields())
        g.step_counter = o.step_counter()
        o.step(1)
        g.step(1)
        (gp, gt, gth), (op, ot, oth) = g.fields(), o.fields()
        assert max_abs(gp, op) <= 1e-6 and max_abs(gt, ot) <= 2e-6, f"after {o.step_counter()} steps"
        d = np.abs(gth.astype(np.float64) - oth.astype(np.float64))
        d = np.minimum(d, np.abs(d - 2 * 3.1415926))
        assert d.max() <= 2e-5
        assert (oth != 0).mean() > 0.9                      # the field really is dense
    if j != int(j):
        return      # eps(theta) jumps at the theta = 0 / 2 pi branch cut for non-integer j: trajectories are not comparable
    o2 = po.Oracle(nx, ny, p, prec=32, math=po.MATH_LIBM, seed=21)
    o2.set_fields(phi, t, z)
    g.set_fields(phi, t, z)
    g.step_counter = 0
    o2.step(40)
    g.step(40)
    assert max_abs(g.phi(), o2.fields()[0]) <= 1e-4 and max_abs(g.t(), o2.fields()[1]) <= 1e-4


def test_fast_far_field_and_symmetry(po, cg):
    g = cg.Kobayashi(256, 256, 1e-4, kernel="fast")
    g.step(5)
    phi, t, th = g.fields()
    assert (phi[:100] == 0).all() and (t[:100] == 0).all() and (th[:100] == 0).all()
    c = 128
    # the single nucleus is mirror symmetric about both axes for the first steps (before rounding chaos)
    assert max_abs(phi[c - 20:c + 21, c - 20:c + 21], phi[c - 20:c + 21, c - 20:c + 21][::-1, :]) <= 1e-6
    assert max_abs(phi[c - 20:c + 21, c - 20:c + 21], phi[c - 20:c + 21, c - 20:c + 21][:, ::-1]) <= 1e-6


@pytest.mark.parametrize("kernel", ["strict", "fast"])
def test_solid_cell_count_long_run(po, cg, kernel):
    """G6: 2000 sub-steps at 250^2: solid-cell count within 2 % of the reference's 10817."""
    g = cg.Kobayashi(250, 250, 1e-4, kernel=kernel)
    for _ in range(200):
        g.iUpdate()
    phi = g.phi()
    solid = int((phi > 0.5).sum())
    assert abs(solid / 10817 - 1) <= 0.02, solid
    assert g.simFrame == 200 and g.simTime > 0
    assert np.isfinite(phi).all() and phi.min() > -0.1 and phi.max() < 1.1


# ---------------------------------------------------------------------------------------------- full size
@pytest.mark.parametrize("kernel", ["strict", "fast"])
def test_4096_single_seed_grid_size_independent_sums(po, cg, kernel):
    """C2 (4096^2, j = 4 and 6, noise off): until the crystal feels the boundary the field sums are independent
    of the grid size, so the 4096^2 GPU run must give the 250^2 oracle sums and exact zeros far away."""
    for j in (4.0, 6.0):
        g = cg.Kobayashi(4096, 4096, 1e-4, kernel=kernel, anisotropy=j)
        g.step(4)
        phi, t, th = g.fields()
        o = po.Oracle(250, 250, po.default_params(anisotropy=j), math=po.MATH_PORTABLE if kernel == "strict" else po.MATH_LIBM)
        o.step(4)
        op, ot, oth = o.fields()
        win = (slice(2048 - 125, 2048 + 125),) * 2
        if kernel == "strict":
            assert bit_equal(phi[win], op) and bit_equal(t[win], ot) and bit_equal(th[win], oth)
        else:
            assert max_abs(phi[win], op) <= 1e-4 and max_abs(t[win], ot) <= 1e-4
        mask = np.ones_like(phi, bool)
        mask[win] = False
        assert (phi[mask] == 0).all() and (t[mask] == 0).all() and (th[mask] == 0).all()


def test_4096_f64_window_vs_oracle(po, cg):
    """C2 FP64: 4096^2 GPU vs 250^2 oracle window, cold 500 sub-steps and warm +500 from the step-500
    state, 1e-10 (BASELINE.json's FP64 tolerance; the strict kernel is in fact bit-equal)."""
    p = po.default_params()
    g = cg.Kobayashi(4096, 4096, 1e-4, precision="f64")
    o = po.Oracle(250, 250, p, prec=64, math=po.MATH_PORTABLE, threads=po.lib().kobo_max_threads())
    x0 = 2048 - 125
    for n in (500, 500):
        g.step(n)
        o.step(n)
        phi, t, _ = g.window(x0, x0, 250, 250)
        assert max_abs(phi, o.fields()[0]) <= 1e-10 and max_abs(t, o.fields()[1]) <= 1e-10
    far = g.window(100, 100, 512, 512)
    assert all((a == 0).all() for a in far)


@pytest.mark.parametrize("kernel,j", [("fast", 6.0), ("fast", 4.0), ("strict", 6.0)])
def test_4096_warm_window_from_oracle_checkpoint(po, cg, kernel, j):
    """C2 FP32 warm window: the oracle's state after 500 sub-steps (250^2, reference arithmetic) is embedded
    in a 4096^2 GPU grid; +1 sub-step stays within 1e-6, +200 within 1e-4 of the reference-arithmetic run (the warm
    trajectory is conditioned well enough for that, SURVEY §5.7: 7.8e-6 under a 1-ulp perturbation), zeros elsewhere."""
    n, w = 4096, 250
    o = po.Oracle(w, w, po.default_params(anisotropy=j), prec=32, math=po.MATH_LIBM, threads=po.lib().kobo_max_threads())
    o.step(500)
    x0, y0 = 1900, 2100                                   # not aligned to anything
    g = cg.Kobayashi(n, n, 1e-4, kernel=kernel, anisotropy=j)
    g.clear()
    g.set_window(x0, y0, *o.fields())
    g.step_counter = o.step_counter()
    o.step(1)
    g.step(1)
    phi, t, th = g.window(x0, y0, w, w)
    assert max_abs(phi, o.fields()[0]) <= 1e-6 and max_abs(t, o.fields()[1]) <= 2e-6
    o.step(199)
    g.step(199)
    phi, t, th = g.window(x0, y0, w, w)
    assert max_abs(phi, o.fields()[0]) <= 1e-4 and max_abs(t, o.fields()[1]) <= 1e-4
    full = g.phi()
    full[y0:y0 + w, x0:x0 + w] = 0
    assert (full == 0).all()                              # the crystal has not left the window; the far field is exactly 0


def _c3_windows(nx, ny, nuclei, w=250, k=3):
    """k nuclei of the bench layout whose w x w window holds no other nucleus and stays inside the grid."""
    out = []
    for (x, y) in nuclei:
        x0, y0 = x - w // 2, y - w // 2
        if x0 < 0 or y0 < 0 or x0 + w > nx or y0 + w > ny:
            continue
        if sum(1 for (a, b) in nuclei if x0 - 4 <= a < x0 + w + 4 and y0 - 4 <= b < y0 + w + 4) == 1:
            out.append((x, y, x0, y0))
        if len(out) == k:
            break
    assert len(out) == k
    return out


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kernel,steps", [("fast", 6), ("strict", 20)])
def test_c3_16384_multi_seed_noise_windows(po, cg, kernel, steps):
    """BASELINE configs[2] AT SIZE: 16384^2, FP32, the bench's 64 Philox-placed nuclei, Philox noise a = 0.01.  Three 250^2
    windows around nuclei are compared with a 250^2 oracle run whose noise keys are the window's GLOBAL coordinates:
    FAST (the default path: two-step launch pairs) for 6 cold sub-steps at BASELINE.json's 1e-4 — a cold FP32 nucleus is
    rounding-chaotic from sub-step 7 — then ONE sub-step from the oracle's own state written into the windows at
    1e-6 (gate G1 at size); STRICT for 20 sub-steps BITWISE; outside the nuclei's neighbourhoods phi is exactly 0."""
    from crystalgrowth_b200.strips import nuclei_positions
    n, w, seed = 16384, 250, 20260101
    nuclei = nuclei_positions(64, n, n, seed)
    g = cg.Kobayashi(n, n, 1e-4, kernel=kernel, seed=seed, noise_a=0.01)
    g.clear()
    for (x, y) in nuclei:
        g.add_nucleus(x, y)
    g.step(steps)
    oracles = []
    for (x, y, x0, y0) in _c3_windows(n, n, nuclei, w):
        o = po.Oracle(w, w, po.default_params(noise_a=0.01), prec=32, seed=seed,
                      math=po.MATH_PORTABLE if kernel == "strict" else po.MATH_LIBM, threads=po.lib().kobo_max_threads())
        o.clear()
        o.set_noise_origin(x0, y0)
        o.add_nucleus(x - x0, y - y0)
        o.step(steps)
        got, want = g.window(x0, y0, w, w), o.fields()
        if kernel == "strict":
            assert all(bit_equal(a, b) for a, b in zip(got, want))
        else:
            assert max_abs(got[0], want[0]) <= 1e-4 and max_abs(got[1], want[1]) <= 1e-4
        oracles.append((o, x0, y0))
    phi = g.phi()
    assert np.isfinite(phi).all()
    near = np.zeros((n, n), bool)
    r = steps * 2 + 4                                     # the composed stencil moves phi by at most 2 cells per sub-step
    for (x, y) in nuclei:
        near[max(y - r, 0):y + r + 1, max(x - r, 0):x + r + 1] = True
    assert (phi[~near] == 0).all() and (phi[near] != 0).sum() > 64 * 5
    del phi, near
    if kernel == "fast":                                  # G1 at size: one sub-step from identical states
        for (o, x0, y0) in oracles:
            g.set_window(x0, y0, *o.fields())
        for (o, x0, y0) in oracles:
            o.step(1)
        g.step(1)
        for (o, x0, y0) in oracles:
            got, want = g.window(x0, y0, w, w), o.fields()
            assert max_abs(got[0], want[0]) <= 1e-6 and max_abs(got[1], want[1]) <= 2e-6


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kernel", ["fast", "strict"])
def test_grid_beyond_2_pow_31_cells(po, cg, kernel):
    """65536 x 36000 = 2.36e9 cells: the reference's `int` cell index (src/Kobayashi.h:91) overflows at 2^31, the new
    layout uses 64-bit offsets (SURVEY §7 hard part 5).  A nucleus at row 35000 (linear index 2.29e9) plus one on the
    x seam; after 4 sub-steps the windows equal a 250^2 oracle run (STRICT bitwise, FAST 1e-6) and rows far away are 0."""
    nx, ny, w = 65536, 36000, 250
    assert nx * ny > 2 ** 31
    g = cg.Kobayashi(nx, ny, 1e-4, kernel=kernel)
    g.clear()
    spots = [(40000, 35000), (65530, 34000)]
    for (x, y) in spots:
        g.add_nucleus(x, y)
    g.step(4)
    o = po.Oracle(w, w, po.default_params(), prec=32, math=po.MATH_PORTABLE if kernel == "strict" else po.MATH_LIBM)
    o.step(4)                                             # reset() seeds the window centre (125, 125)
    want = o.fields()
    got = g.window(40000 - 125, 35000 - 125, w, w)
    if kernel == "strict":
        assert all(bit_equal(a, b) for a, b in zip(got, want))
    else:
        assert max_abs(got[0], want[0]) <= 1e-6 and max_abs(got[1], want[1]) <= 2e-6
    # the nucleus on the seam: its right half wraps to columns 0..; compare the two halves with the oracle window
    left = g.window(nx - 125, 34000 - 125, 125, w)        # window columns 0..124  <- global 65411..65535
    right = g.window(0, 34000 - 125, 125, w)              # window columns 125..249 <- global 0..124
    o2 = po.Oracle(w, w, po.default_params(), prec=32, math=po.MATH_PORTABLE if kernel == "strict" else po.MATH_LIBM)
    o2.clear()
    o2.add_nucleus(65530 - (nx - 125), 125)
    o2.step(4)
    for k in range(3):
        whole = np.concatenate([left[k], right[k]], axis=1)
        if kernel == "strict":
            assert bit_equal(whole, o2.fields()[k])
        elif k < 2:
            assert max_abs(whole, o2.fields()[k]) <= 2e-6
    for y0 in (0, 17000, 33000):
        assert all((a == 0).all() for a in g.window(1000, y0, 2048, 64))


# ---------------------------------------------------------------------------------------------- API behaviour
def test_roundtrip_reset_params_counters(po, cg):
    g = cg.Kobayashi(70, 50, 1e-4, kernel="strict")
    rng = np.random.default_rng(0)
    phi = rng.random((50, 70), np.float32)
    t = rng.random((50, 70), np.float32)
    th = rng.random((50, 70), np.float32)
    g.set_fields(phi, t, th)
    a, b, c = g.fields()
    assert bit_equal(a, phi) and bit_equal(b, t) and bit_equal(c, th)
    g.set_fields(phi=None, t=None, angl=None)
    g.reset()
    a, b, c = g.fields()
    assert a.sum() == 5 and a[25, 35] == 1 and a[25, 34] == 1 and a[24, 35] == 1 and b.sum() == 0 and c.sum() == 0
    g.set_params(K=1.3, reset=True)
    assert g.K == 1.3 and g.get_params().tau == 0.0003
    n0 = g.launch_count
    g.step(7)
    g.sync()
    assert g.launch_count == n0 + 7 and g.step_counter == 7
    g.iUpdate()
    assert g.simFrame == 1 and g.step_counter == 17
    g.reset()
    assert g.simFrame == 0 and g.step_counter == 0


def test_nucleus_wraps_like_the_oracle(po, cg):
    o, g = _pair(po, cg, 16, 12, nuclei=[(0, 0), (15, 11), (-1, 5), (7, 12)])
    _assert_bitwise(o, g)
    o.step(3)
    g.step(3)
    _assert_bitwise(o, g)


def test_errors_are_reported_not_thrown(po, cg):
    with pytest.raises(cg.KobayashiError) as e:
        cg.Kobayashi(1, 10, 1e-4)
    assert e.value.status == -1
    with pytest.raises(cg.KobayashiError):
        cg.Kobayashi(32, 32, 1e-4, tau=0.0)
    with pytest.raises(cg.KobayashiError) as e:
        cg.Kobayashi(32, 32, 1e-4, precision="f64", kernel="fast")
    assert e.value.status == -6
    with pytest.raises(cg.KobayashiError):
        cg.Kobayashi(32, 32, 1e-4, device=99)
    g = cg.Kobayashi(32, 32, 1e-4, kernel="strict")
    with pytest.raises(cg.KobayashiError):
        g.step(-1)
    with pytest.raises(ValueError):
        g.set_fields(np.zeros((3, 3), np.float32))


def test_window_access_and_async_readback(po, cg):
    """kob_get_window / kob_set_window (sub-rectangles; periodic aliases of written seam cells refreshed), kob_get_fields_async +
    kob_wait_fields (snapshot, then stepping continues while the copy runs), kob_host_alloc_near (pinned, NUMA-near)."""
    import ctypes as C
    for kernel in ("fast", "strict"):
        nx, ny = 300, 200
        g = cg.Kobayashi(nx, ny, 1e-4, kernel=kernel, noise_a=0.01, seed=4)
        g.step(30)
        phi, t, th = g.fields()
        for (x0, y0, w, h) in [(0, 0, 7, 5), (120, 80, 64, 40), (nx - 9, ny - 6, 9, 6), (0, 0, nx, ny)]:
            a, b, c_ = g.window(x0, y0, w, h)
            assert bit_equal(a, phi[y0:y0 + h, x0:x0 + w]) and bit_equal(b, t[y0:y0 + h, x0:x0 + w]) and bit_equal(c_, th[y0:y0 + h, x0:x0 + w])
        with pytest.raises(cg.KobayashiError):
            g.window(nx - 3, 0, 8, 4)
        # a window written over the x seam and the y seam must step like the same state written whole
        rng = np.random.default_rng(3)
        patch = (0.5 + 0.4 * rng.random((6, 9))).astype(np.float32)
        g2 = cg.Kobayashi(nx, ny, 1e-4, kernel=kernel, noise_a=0.01, seed=4)
        g2.set_fields(phi, t, th)
        g2.step_counter = g.step_counter
        phi2 = phi.copy()
        phi2[ny - 6:, nx - 9:] = patch
        g.set_window(nx - 9, ny - 6, phi=patch)
        g2.set_fields(phi2, t, th)
        g.step(3)
        g2.step(3)
        assert all(bit_equal(x, y) for x, y in zip(g.fields(), g2.fields()))
        # asynchronous readback: the snapshot is the state at the call, whatever is queued afterwards
        n = nx * ny * 4
        bufs = [g.host_alloc_near(n) for _ in range(3)]
        want = g.fields()
        g.get_fields_async(bufs[0], bufs[1], bufs[2])
        g.step(5)
        g.wait_fields()
        for k in range(3):
            got = np.frombuffer((C.c_char * n).from_address(bufs[k].value), np.float32).reshape(ny, nx)
            assert bit_equal(got.copy(), want[k])
        L = cg.load()
        for p_ in bufs:
            assert L.kob_host_free(p_) == 0
        g.close()
        g2.close()


def test_ring_guards(cg):
    """Linked strips must run one launch sequence: thin FAST strips are refused, a host-injected noise field needs the
    ring in single-step mode, kob_ring_join checks its arguments."""
    a = cg.Kobayashi(64, 6, 1e-4, kernel="fast", ny_global=38, y0=0)
    b = cg.Kobayashi(64, 32, 1e-4, kernel="fast", ny_global=38, y0=6)
    with pytest.raises(cg.KobayashiError) as e:
        a.link_local(b, b)
    assert "8 rows" in str(e.value)
    s0 = cg.Kobayashi(64, 32, 1e-4, kernel="fast", ny_global=64, y0=0, noise_a=0.01)
    s1 = cg.Kobayashi(64, 32, 1e-4, kernel="fast", ny_global=64, y0=32, noise_a=0.01)
    s0.link_local(s1, s1)
    s1.link_local(s0, s0)
    r = np.full((32, 64), 0.25, np.float32)
    with pytest.raises(cg.KobayashiError):
        s0.set_noise_field(r)
    for s in (s0, s1):
        s.set_path_mode(0)
    s0.set_noise_field(r)
    s1.set_noise_field(r)
    for s in (s0, s1):
        s.halo_refresh()
    for s in (s0, s1):
        s.step(4)
    for s in (s0, s1):
        s.sync()
    with pytest.raises(cg.KobayashiError):
        s0.ring_join("x" * 60, 0, 2)
    with pytest.raises(cg.KobayashiError):
        s0.ring_join("ok", 2, 2)
    for s in (a, b, s0, s1):
        s.close()


def test_render_matches_reference_colours(po, cg):
    """kob_render_rgba vs the colours the REFERENCE gives every object (iUpdateConstantBuffer, src/Kobayashi.cpp:309-345):
    the committed fixture tests/golden/ref_colors_n48.npz was produced by the reference TU (scripts/make_golden.py,
    run_colors) for a designed phi field incl. the ramp's break points; where oracle/_ref is available the reference is
    also run live on a second field.  Object i shows cell (x, y) = (i / n, i % n) (:311-315); the image is [y, x]."""
    z = np.load(os.path.join(GOLDEN, "ref_colors_n48.npz"))
    n, phi, rgb = int(z["n"]), z["phi"], z["rgb"]
    cases = [(phi, rgb)]
    if po.ref_available(32):
        rng = np.random.default_rng(5)
        phi2 = (rng.random((n, n)) * 1.2 - 0.1).astype(np.float32)
        r = po.Reference(n, n, 1e-4, prec=32)
        r.set_fields(phi2, np.zeros_like(phi2), np.zeros_like(phi2))
        cases.append((phi2, r.colors()))
    i = np.arange(n * n)
    for kernel in ("strict", "fast"):
        g = cg.Kobayashi(n, n, 1e-4, kernel=kernel)
        for p_, want in cases:
            g.set_fields(p_, None, None)
            img = g.render_rgba()
            got = img[i % n, i // n, :3].astype(np.int32)                  # object i -> pixel (x, y) = (i / n, i % n)
            exp = np.rint(np.clip(want.astype(np.float64), 0, 1) * 255).astype(np.int32)
            assert np.abs(got - exp).max() <= 1
            assert (img[..., 3] == 255).all()
        g.close()


# ---------------------------------------------------------------------------------------------- strips on one GPU
@pytest.mark.timeout(300)
@pytest.mark.parametrize("kernel,prec", [("strict", "f32"), ("fast", "f32"), ("strict", "f64")])
@pytest.mark.parametrize("nstrips,nyg", [(2, 64), (3, 70), (4, 37)])
def test_linked_strips_on_one_gpu_equal_single_domain(po, cg, kernel, prec, nstrips, nyg):
    """G5 (shard invariance): P row strips linked with kob_link_local on ONE device — the same edge-tile peer
    stores and step flags the multi-GPU ring uses — reproduce the single-domain run bit-for-bit."""
    from crystalgrowth_b200.strips import partition
    nx, steps, seed = 96, 40, 5
    kw = dict(precision=prec, kernel=kernel, seed=seed, noise_a=0.01)
    nuclei = [(3, 0), (48, nyg // 2), (95, nyg - 1), (20, nyg // nstrips), (70, nyg // nstrips - 1)]
    single = cg.Kobayashi(nx, nyg, 1e-4, **kw)
    single.clear()
    strips = [cg.Kobayashi(nx, ny, 1e-4, ny_global=nyg, y0=y0, **kw) for (y0, ny) in partition(nyg, nstrips)]
    for i, s in enumerate(strips):
        s.link_local(strips[(i - 1) % nstrips], strips[(i + 1) % nstrips])
        s.clear()
    for (x, y) in nuclei:
        single.add_nucleus(x, y)
        for s in strips:
            s.add_nucleus(x, y)
    for s in strips:
        s.sync()
    for s in strips:
        s.halo_refresh()
    for s in strips:
        s.sync()
    for _ in range(steps):            # lock-step issue order: step e of every strip precedes step e+1 of any
        for s in strips:
            s.step(1)
    single.step(steps)
    want = single.fields()
    got = [np.concatenate(parts, axis=0) for parts in zip(*[s.fields() for s in strips])]
    for a, b in zip(got, want):
        assert bit_equal(a, b)
    for s in strips:
        s.close()


# ---------------------------------------------------------------------------------------------- FAST variants
@pytest.mark.parametrize("cta", [0, 1, 2])
@pytest.mark.parametrize("free", [0, 1])
def test_fast_far_field_shortcut_is_bit_neutral(cg, monkeypatch, cta, free):
    """Chunks whose phi rows (and the 4 rows before them) are all +0 skip the phi arithmetic and only diffuse T.
    That must not change a single bit: run with the shortcut disabled (KOB_FAST_NOSKIP=1) and compare.  The grid
    is wide/tall enough for whole far-field jobs, partial ones next to the crystals, and heat (T != 0) diffusing
    into phi == 0 territory; noise is on."""
    def run(noskip):
        monkeypatch.setenv("KOB_FAST_NOSKIP", str(noskip))
        monkeypatch.setenv("KOB_FAST2", "0")             # the single-step kernel's shortcut is what is under test
        monkeypatch.setenv("KOB_FAST_CTA", str(cta))
        monkeypatch.setenv("KOB_FAST_FREE", str(free))    # CTAs that met a crystal switch to per-warp job claims, or never do
        monkeypatch.setenv("KOB_FAST_YJ", "32")
        g = cg.Kobayashi(700, 300, 1e-4, kernel="fast", seed=11, noise_a=0.01)
        g.clear()
        for (x, y) in [(0, 0), (350, 150), (699, 299), (100, 40), (520, 222)]:
            g.add_nucleus(x, y)
        g.step(150)
        out = g.fields()
        g.close()
        return out
    a, b = run(0), run(1)
    assert (a[0] == 0).mean() > 0.5 and (a[1] != 0).mean() > (a[0] != 0).mean()      # the case is really exercised
    assert all(bit_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("cta", [0, 1, 2])
@pytest.mark.parametrize("np_,yj", [(1, 8), (0, 256), (0, 12), (1, 5)])
def test_fast_tuning_variants(po, cg, monkeypatch, np_, yj, cta):
    """The FAST kernel's decomposition knobs (per-warp job claims after a CTA met a crystal: on / off; rows per job) do not
    change results: every variant passes the single-step gate, and the knobs are bit-neutral against the default 256-row run."""
    def run(env_np, env_yj, nx=150, ny=90, steps=12):
        monkeypatch.setenv("KOB_FAST_FREE", str(env_np))
        monkeypatch.setenv("KOB_FAST_YJ", str(env_yj))
        monkeypatch.setenv("KOB_FAST_CTA", str(cta))
        g = cg.Kobayashi(nx, ny, 1e-4, kernel="fast", seed=9, noise_a=0.01)
        g.clear()
        for (x, y) in [(0, 0), (75, 45), (149, 89), (30, 7), (120, 8)]:
            g.add_nucleus(x, y)
        g.step(steps)
        return g
    a = run(np_, yj).fields()
    b = run(1, 256).fields()
    assert all(bit_equal(x, y) for x, y in zip(a, b))
    # single-step gate from an evolved oracle state
    o = po.Oracle(150, 90, po.default_params(noise_a=0.01), math=po.MATH_LIBM, seed=9)
    o.clear()
    for (x, y) in [(0, 0), (75, 45), (149, 89), (30, 7), (120, 8)]:
        o.add_nucleus(x, y)
    o.step(30)
    monkeypatch.setenv("KOB_FAST_FREE", str(np_))
    monkeypatch.setenv("KOB_FAST_YJ", str(yj))
    g = cg.Kobayashi(150, 90, 1e-4, kernel="fast", seed=9, noise_a=0.01)
    g.set_fields(*o.fields())
    g.step_counter = o.step_counter()
    o.step(1)
    g.step(1)
    assert max_abs(g.phi(), o.fields()[0]) <= 1e-6 and max_abs(g.t(), o.fields()[1]) <= 2e-6


def test_launch_trace_file(cg, monkeypatch, tmp_path):
    """KOB_TRACE=<file>: CUDA events around the step kernels' launches, written at kob_destroy (diagnostics; DESIGN §4.1b's numbers
    come from it).  One record per launch, or per launch pair when the general pass runs beside the far pass."""
    path = tmp_path / "trace_%p.csv"
    monkeypatch.setenv("KOB_TRACE", str(path))
    monkeypatch.setenv("KOB_FAST2", "1")
    g = cg.Kobayashi(600, 400, 1e-4, kernel="fast", seed=3, noise_a=0.01)
    g.step(20)
    g.sync()
    g.close()
    import os
    files = [f for f in os.listdir(tmp_path) if f.startswith("trace_") and f.endswith(".csv")]
    assert len(files) == 1 and "%p" not in files[0]
    lines = open(tmp_path / files[0]).read().strip().splitlines()
    assert lines[0] == "kernel,start_us,duration_us,gap_before_us,info"
    kinds = [ln.split(",")[0] for ln in lines[1:]]
    pairs = sum(k.startswith("pair") for k in kinds) + sum(k == "kob_far2" for k in kinds)
    assert pairs == 10 and all(float(ln.split(",")[2]) > 0 for ln in lines[1:])
