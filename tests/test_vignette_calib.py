"""vignetteCalib's alternating optimiser (main_vignetteCalib.cpp:395-585): the oracle restatement on its
own (CPU) and the CUDA kernels against it (GPU)."""
import numpy as np
import pytest

from conftest import assert_bits_equal
from mono_dataset_code_b200 import synthetic as S
from oracle.loader import PortOracle

GW, GH, WI, HI, N = 96, 80, 120, 90, 14


@pytest.fixture(scope="module")
def port():
    return PortOracle()


@pytest.fixture(scope="module")
def problem():
    return S.vignette_calib_problem(N, GW, GH, WI, HI, seed=3)


def oracle_loop(port, pr, iters, outlier_th, plane0, vig0, int_abs=True):
    plane, vig, log = plane0.copy(), vig0.copy(), []
    for it in range(iters):
        oth2 = outlier_th * outlier_th if it >= iters // 2 else 10000 * 10000
        plane, _, _, ps = port.vc_plane_step(pr["images"], pr["p2x"], pr["p2y"], WI, HI, vig, plane, oth2, int_abs)
        vig, _, _, vs = port.vc_vignette_step(pr["images"], pr["p2x"], pr["p2y"], WI, HI, plane, vig, oth2, int_abs)
        log.append([ps[0], ps[1], vs[0], vs[1]])
    return plane, vig, np.array(log)


def test_oracle_loop_recovers_the_synthetic_vignette(port, problem):
    """Sanity of the restatement: on a rendered scene with a known vignette the loop converges towards it."""
    plane, vig, log = oracle_loop(port, problem, 8, 15, np.zeros(GW * GH, np.float32), np.ones(WI * HI, np.float32))
    truth = problem["true_vignette"] / np.nanmax(problem["true_vignette"])
    m = np.isfinite(vig)
    assert m.sum() > 0.5 * WI * HI
    err = np.abs(vig[m] / np.nanmax(vig) - truth[m])
    assert np.median(err) < 0.03 and np.percentile(err, 95) < 0.12
    rms = np.sqrt(log[:, 2] / log[:, 3])
    assert rms[-1] < rms[0]                      # the residual of the vignette step goes down
    assert np.all(log[:, 1] > 0) and np.all(log[:, 3] > 0)


def test_oracle_smoothing_fills_nans_and_keeps_constants(port):
    v = np.full((HI, WI), 0.5, np.float32)
    v[10:14, 20:25] = np.nan
    v[0, 0] = np.nan
    out = port.vc_smooth(v.ravel(), WI, HI, 4).reshape(HI, WI)
    assert np.all(np.isfinite(out)) and np.allclose(out, 0.5)
    allnan = np.full(WI * HI, np.nan, np.float32)
    assert np.all(np.isnan(port.vc_smooth(allnan, WI, HI, 2)))


# ------------------------------------------------------------------------------------------------ GPU parity
@pytest.fixture(scope="module")
def gpu(problem):
    torch = pytest.importorskip("torch")
    from mono_dataset_code_b200 import api
    ctx = api.Context(None, None, 0)
    d = {k: torch.from_numpy(problem[k]).cuda() for k in ("images", "p2x", "p2y")}
    return torch, ctx, d


def start_state(seed=5):
    rng = np.random.default_rng(seed)
    plane = rng.uniform(60, 180, GW * GH).astype(np.float32)
    plane[rng.integers(0, GW * GH, 40)] = np.nan
    vig = rng.uniform(0.5, 1.0, WI * HI).astype(np.float32)
    vig[rng.integers(0, WI * HI, 60)] = np.nan
    return plane, vig


# integer_abs: `abs(residual) > oth2` with the residual truncated to int (the reference's era, and the reference program as built
# for the oracle) or fabs (what the same line means with <cmath>'s overloads in scope) -- include/mdc_b200.h
MODES = pytest.mark.parametrize("oth2,int_abs", [(10000 * 10000, True), (400, True), (400, False), (9, True), (9, False)],
                                ids=["no_outliers", "th20_int_abs", "th20_fabs", "th3_int_abs", "th3_fabs"])


@pytest.mark.gpu
@MODES
def test_plane_step_is_bit_identical(port, problem, gpu, oth2, int_abs):
    torch, ctx, d = gpu
    plane0, vig0 = start_state()
    exp, _, _, st = port.vc_plane_step(problem["images"], problem["p2x"], problem["p2y"], WI, HI, vig0, plane0, oth2, int_abs)
    pc = torch.from_numpy(plane0).cuda()
    E, R = ctx.vc_plane_step(d["images"], d["p2x"], d["p2y"], GW, GH, WI, HI, torch.from_numpy(vig0).cuda(), pc, oth2, int_abs)
    assert_bits_equal(pc.cpu().numpy(), exp, "plane colour")
    assert R == st[1] and abs(E - st[0]) <= 1e-9 * abs(st[0])


@pytest.mark.gpu
@MODES
def test_vignette_step_matches_to_rounding(port, problem, gpu, oth2, int_abs):
    """The reference adds the scattered terms in (image, point) order in fp32; the GPU adds the same terms exactly (64-bit fixed
    point), so it differs from the reference by the rounding error of the reference's chain only — and gives the same bits on
    every run."""
    TOL = 2e-5
    torch, ctx, d = gpu
    plane0, vig0 = start_state()
    exp, tt, _, st = port.vc_vignette_step(problem["images"], problem["p2x"], problem["p2y"], WI, HI, plane0, vig0, oth2, int_abs)
    v = torch.from_numpy(vig0).cuda()
    E, R = ctx.vc_vignette_step(d["images"], d["p2x"], d["p2y"], GW, GH, WI, HI, torch.from_numpy(plane0).cuda(), v, oth2, int_abs)
    got = v.cpu().numpy()
    firm = np.abs(tt - 1.0) > 1e-3              # pixels whose "TT < 1" test cannot flip with the summation order
    assert np.array_equal(np.isnan(got)[firm], np.isnan(exp)[firm])
    m = np.isfinite(exp) & np.isfinite(got)
    assert m.sum() > 200
    assert np.max(np.abs(got[m] - exp[m]) / np.maximum(np.abs(exp[m]), 1e-6)) < TOL
    assert R == st[1] and abs(E - st[0]) <= 1e-9 * abs(st[0])
    for _ in range(2):                           # order-independent sums: repeated runs agree bit for bit (values and statistics)
        v2 = torch.from_numpy(vig0).cuda()
        E2, R2 = ctx.vc_vignette_step(d["images"], d["p2x"], d["p2y"], GW, GH, WI, HI, torch.from_numpy(plane0).cuda(), v2, oth2, int_abs)
        assert_bits_equal(v2.cpu().numpy(), got, "vignette step, repeated")
        assert (E2, R2) == (E, R)
    # the sums are the exact sums of the reference's fp32 terms: compare TT-weighted against a float64 accumulation of the oracle's terms
    if oth2 == 10000 * 10000:
        exact = exact_vignette_step(problem, plane0, vig0)
        m = np.isfinite(exact) & np.isfinite(got) & firm
        assert np.max(np.abs(got[m] - exact[m]) / np.maximum(np.abs(exact[m]), 1e-6)) < 6e-7      # a few fp32 ulps: two conversions, one division, the normalisation


def exact_vignette_step(pr, plane, vig):
    """The vignette step with the reference's fp32 terms but float64 sums (numpy), no outlier rejection: what the fixed-point
    accumulation must reproduce up to the final float conversions."""
    f32 = np.float32
    tt, ct = np.zeros(WI * HI), np.zeros(WI * HI)
    for img in range(N):
        x, y = pr["p2x"][img], pr["p2y"][img]
        ok = ~np.isnan(x) & ~np.isnan(plane)
        xi, yi = x[ok], y[ok]
        ix, iy = xi.astype(np.int32), yi.astype(np.int32)
        dx, dy = (xi - ix.astype(f32)).astype(f32), (yi - iy.astype(f32)).astype(f32)
        dxdy = (dx * dy).astype(f32)
        w11, w01, w10 = dxdy, (dy - dxdy).astype(f32), (dx - dxdy).astype(f32)
        w00 = (((f32(1.0) - dx).astype(f32) - dy).astype(f32) + dxdy).astype(f32)
        base = ix + iy * WI
        im = pr["images"][img]
        cI = ((((w11 * im[base + 1 + WI]).astype(f32) + (w01 * im[base + WI]).astype(f32)).astype(f32) + (w10 * im[base + 1]).astype(f32)).astype(f32)
              + (w00 * im[base]).astype(f32)).astype(f32)
        cP = plane[ok]
        good = ~np.isnan(cI)
        for w, off in ((w00, 0), (w10, 1), (w01, WI), (w11, 1 + WI)):
            np.add.at(tt, base[good] + off, ((w[good] * cP[good]).astype(f32) * cP[good]).astype(f32).astype(np.float64))
            np.add.at(ct, base[good] + off, ((w[good] * cI[good]).astype(f32) * cP[good]).astype(f32).astype(np.float64))
    t32, c32 = tt.astype(f32), ct.astype(f32)
    out = np.full(WI * HI, np.nan, f32)
    m = ~(t32 < 1)
    out[m] = c32[m] / t32[m]
    return (out / np.nanmax(out)).astype(f32)


@pytest.mark.gpu
def test_smoothing_is_bit_identical(port, gpu):
    torch, ctx, _ = gpu
    rng = np.random.default_rng(9)
    v = rng.uniform(0.2, 1.0, WI * HI).astype(np.float32)
    v[rng.integers(0, WI * HI, 900)] = np.nan
    v.reshape(HI, WI)[30:50, 40:70] = np.nan          # a hole wider than 4 rounds can close
    for iters in (0, 1, 4):
        exp = port.vc_smooth(v, WI, HI, iters)
        got = ctx.vc_smooth(torch.from_numpy(v).cuda(), WI, HI, iters).cpu().numpy()
        assert_bits_equal(got, exp, f"smoothing x{iters}")


@pytest.mark.gpu
@pytest.mark.parametrize("int_abs", [True, False], ids=["int_abs", "fabs"])
def test_whole_loop_against_the_oracle_loop(port, problem, gpu, capfd, int_abs):
    torch, ctx, d = gpu
    iters, th = 6, 15
    plane0, vig0 = np.zeros(GW * GH, np.float32), np.ones(WI * HI, np.float32)
    plane_e, vig_e, log_e = oracle_loop(port, problem, iters, th, plane0, vig0, int_abs)
    pc, v = torch.from_numpy(plane0).cuda(), torch.from_numpy(vig0).cuda()
    smoothed, log = ctx.vignette_calib(d["images"], d["p2x"], d["p2y"], GW, GH, WI, HI, iters, th, pc, v, int_abs)
    assert "residual terms =>" in capfd.readouterr().out          # the reference's progress lines (:448, :519)
    got_v, got_p = v.cpu().numpy(), pc.cpu().numpy()
    m = np.isfinite(vig_e) & np.isfinite(got_v)
    assert m.sum() > 0.5 * WI * HI and np.mean(np.isnan(got_v) == np.isnan(vig_e)) > 0.999
    assert np.max(np.abs(got_v[m] - vig_e[m])) < 2e-4
    mp = np.isfinite(plane_e) & np.isfinite(got_p)
    assert np.max(np.abs(got_p[mp] - plane_e[mp]) / np.maximum(np.abs(plane_e[mp]), 1.0)) < 2e-4
    assert np.allclose(log[:, 1], log_e[:, 1], rtol=2e-3) and np.allclose(log[:, 0], log_e[:, 0], rtol=2e-3)
    s_exp = port.vc_smooth(got_v, WI, HI, 4)
    assert_bits_equal(smoothed.cpu().numpy(), s_exp, "smoothed output")
