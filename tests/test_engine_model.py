"""CPU tests of the FFT engine's math (numpy model) and of the C planner behind the C ABI."""
import numpy as np
import pytest

import engine_model as em

LENGTHS = [1024, 2048, 4096, 8192, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 19, 21, 30, 35, 46, 64, 66, 106, 130, 132, 148, 158, 168,
           218, 256, 286, 300, 346, 384, 512, 560, 18, 36, 45, 75, 90, 96, 135, 150, 210, 270, 420, 448, 1080, 1125, 2160]


def test_composite_radix_plans():
    want = {1125: [15, 15, 5], 2160: [16, 15, 9], 1080: [8, 15, 9], 810: [6, 15, 9], 540: [4, 15, 9],
            # y / z axes: at most three stages where small primes pair into composite radices
            150: [10, 15], 135: [15, 9], 210: [2, 15, 7], 360: [8, 15, 3], 480: [8, 4, 15], 288: [8, 4, 9], 350: [10, 5, 7],
            600: [8, 15, 5], 96: [8, 4, 3], 384: [16, 8, 3], 192: [8, 8, 3], 560: [16, 5, 7], 280: [8, 5, 7], 224: [8, 4, 7],
            400: [16, 5, 5],
            # two stages of fat composite radices where they were measured to win (y axis)
            270: [18, 15], 300: [20, 15], 420: [20, 21], 448: [16, 28]}
    for L, r in want.items():
        assert em.factorize(L) == r, L
    # fused z axis / x axis variants
    assert em.factorize(448, 1) == [8, 8, 7] and em.factorize(560, 1) == [28, 20] and em.factorize(300, 1) == [20, 15]
    assert em.factorize(150, 2) == [10, 15] and em.factorize(135, 2) == [9, 15] and em.factorize(300, 2) == [4, 3, 5, 5]
    # x axis: one stage per prime up to four stages (the tiled kernels lose with composite radices)
    assert em.factorize(210, 2) == [2, 3, 5, 7] and em.factorize(360, 2) == [8, 3, 3, 5] and em.factorize(350, 2) == [2, 5, 5, 7]


@pytest.mark.parametrize("L", [2, 3, 4, 5, 7, 8, 12, 30, 35, 64, 66, 130, 158, 300, 270, 420, 1125, 90])
def test_inplace_dif_and_mirror_inverse(L):
    rng = np.random.default_rng(L)
    r = em.factorize(L)
    x = rng.standard_normal(L) + 1j * rng.standard_normal(L)
    y = em.fwd_inplace(x, r)
    assert np.allclose(y, np.fft.fft(x)[em.rev_positions(L, r)])
    assert np.allclose(em.inv_inplace(y, r), L * x)


@pytest.mark.parametrize("n", [2, 4, 6, 10, 16, 30, 64, 70, 130, 512])
def test_packed_r2c_c2r(n):
    rng = np.random.default_rng(n)
    r = em.factorize(n // 2)
    x = rng.standard_normal(n)
    X, P = em.r2c_even(x, r)
    assert np.allclose(np.concatenate([X[P], X[n // 2:]]), np.fft.rfft(x))
    assert np.allclose(em.c2r_even(X, n, r), n * x)


def test_swizzle_is_conflict_free():
    for p0 in range(0, 256, 8):
        assert len({em.swz(p0 + i) for i in range(8)}) == 8
    for a in range(0, 256, 16):
        assert len({em.swz(a + 2 * i) for i in range(8)}) == 8
        assert len({em.swz(a + 2 * i + 1) for i in range(8)}) == 8


@pytest.mark.parametrize("style", [0, 1, 2])
@pytest.mark.parametrize("L", LENGTHS)
def test_c_planner_matches_model(fc, L, style):
    radices, generic = fc.plan_radices(L, style)
    assert radices == em.factorize(L, style)
    assert int(np.prod(radices)) == L
    assert generic == any(r not in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 25, 28)
                          for r in radices)
    rev, pos, tw = fc.plan_tables(L, style)
    assert np.array_equal(rev, em.rev_positions(L, radices))
    assert np.array_equal(pos[rev], np.arange(L))
    assert np.abs(tw - em.twiddles(L)).max() < 1e-7


def test_principal_roots_are_exact(fc):
    _, _, tw = fc.plan_tables(512)
    assert tw[0] == 1 and tw[128] == -1j and tw[256] == -1 and tw[384] == 1j


def test_psf_active_rows_match_oracle_placement(fc):
    from oracle import fc_oracle as fo
    for imDim, kDim in (([16, 16, 16], [3, 3, 3]), ([15, 19, 21], [3, 3, 3]), ([46, 46, 106], [31, 31, 91]),
                        ([20, 12, 10], [5, 4, 3])):
        S = fo.place_psf(np.ones(int(np.prod(kDim)), np.float32), kDim, imDim)
        rows = np.unique(np.nonzero(S)[0] // imDim[0])
        assert np.array_equal(rows, fc.psf_active_rows(imDim, kDim))


def test_c_planner_matches_model_exhaustively_up_to_1200(fc):
    """every length a caller can pass on an axis (up to 1200, plus the padded config-5 extents): same radix
    sequence as the numpy model, product == L, stages bounded, digit reversal a permutation"""
    fast = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 25, 28)     # register butterflies
    smooth7 = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 18, 20, 21, 25, 28)
    for L in list(range(1, 1201)) + [1125, 2048, 2160, 4096]:
        for style in (0, 1, 2):
            radices, generic = fc.plan_radices(L, style)
            assert radices == em.factorize(L, style), (L, style)
            assert int(np.prod(radices)) == L
            assert generic == any(r not in fast for r in radices)
        r0 = em.factorize(L)
        smooth = [r for r in r0 if r in smooth7]
        if len(smooth) == len(r0) and L > 1:
            # 7-smooth lengths never need more than four shared-memory round trips up to 1200 ... except the few with
            # many repeated small primes that even composite radices cannot pack into four stages
            assert len(r0) <= 5, (L, r0)
    rev, pos, _ = fc.plan_tables(270)
    assert sorted(rev.tolist()) == list(range(270)) and np.array_equal(pos[rev], np.arange(270))


@pytest.mark.parametrize("p", [37, 53, 79, 109, 271, 281, 163, 61])
def test_rader_tables_reproduce_the_dft(fc, p):
    """numpy walk through the device Rader stage (fft_engine.cuh: stage_rader) with the tables the C planner builds:
    gather a[m] = x[g^m], n-point DIF (position order), multiply by the stored spectrum of b (x[0] joins the DC term,
    X[0] = x[0] + A[0]), unnormalised n-point inverse, scatter to g^-q -- must equal numpy's DFT, forward and inverse"""
    r = fc.plan_rader(p)
    assert r is not None, p
    n = p - 1
    assert int(np.prod(r["radices"])) == n
    radices, generic = fc.plan_radices(n, 3)    # style 3: the sub-transform never uses the two-stage fat plans
    assert radices == r["radices"] and not generic
    rev, pos, _ = fc.plan_tables(n, 3)
    assert sorted(r["perm"].tolist()) == list(range(1, p)) and sorted(r["iperm"].tolist()) == list(range(1, p))
    assert all((int(a) * int(b)) % p == 1 for a, b in zip(r["perm"], r["iperm"]))
    rng = np.random.default_rng(p)
    x = rng.standard_normal(p) + 1j * rng.standard_normal(p)
    for B, want in ((r["bf"], np.fft.fft(x)), (r["bi"], np.fft.ifft(x) * p)):
        a = x[r["perm"]]
        A = np.fft.fft(a)
        Apos = A[rev]                       # position q holds frequency rev[q]
        Cpos = Apos * B
        X0 = x[0] + Apos[0]
        Cpos[0] += x[0]
        C = np.zeros(n, complex)
        C[rev] = Cpos
        c = np.fft.ifft(C) * n              # unnormalised inverse (1/n is folded into B)
        X = np.zeros(p, complex)
        X[0] = X0
        X[r["iperm"]] = c
        assert np.abs(X - want).max() <= 2e-5 * np.abs(want).max()


def test_rader_is_declined_when_p_minus_1_is_not_smooth(fc):
    assert fc.plan_rader(173) is None       # 172 = 4 * 43
    assert fc.plan_rader(47) is None        # 46 = 2 * 23: radix 23 is not among the n-point transform's radices


def test_plan_override_parsing(fc, monkeypatch):
    """FCB200_PLAN (tuning runs, fc_plan.cu: plan_override): '<axis><length>=r0.r1...' entries, comma separated; an entry
    applies to its axis style and length only, products that do not match and unknown axes are ignored"""
    monkeypatch.setenv("FCB200_PLAN", "z300=15.20,y420=12.5.7,x280=20.14,y512=8.8,q64=8.8,y96=")
    assert fc.plan_radices(300, 1) == ([15, 20], False)
    assert fc.plan_radices(300, 0) == ([20, 15], False)            # other axis: the planner's own
    assert fc.plan_radices(420, 0) == ([12, 5, 7], False)
    assert fc.plan_radices(280, 2) == ([20, 14], False)
    assert fc.plan_radices(512, 0) == ([8, 8, 8], False)           # 8 * 8 != 512: ignored
    assert fc.plan_radices(64, 0) == ([8, 8], False) and fc.plan_radices(96, 0) == ([8, 4, 3], False)
    monkeypatch.setenv("FCB200_PLAN", "y158=2.79")
    assert fc.plan_radices(158, 0) == ([2, 79], True)              # a prime above 23 still means the direct-sum stage
    rev, pos, _ = fc.plan_tables(158, 0)
    assert np.array_equal(pos[rev], np.arange(158))
    monkeypatch.delenv("FCB200_PLAN")
    assert fc.plan_radices(300, 1) == ([20, 15], False)
