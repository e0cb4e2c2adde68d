"""CPU checks of the geometric pre-processing row: the numpy oracle against the committed golden
vectors (written by tools/make_golden_remap.py from cv2.remap and the unmodified reference functions), against the live
cv2 when it is installed, and the host-side logic of the product (map builders, weight table, max-rectangle scan)
against the oracle.  No GPU work: the bit-exact CUDA-vs-oracle tests are in test_preprocess_gpu.py."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import remap as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "remap_golden.npz"))


def test_oracle_remap_matches_cv2_golden():
    out = R.remap_cubic_wrap_u8(G["raw_img"], G["raw_mx"], G["raw_my"])
    assert np.array_equal(out, G["raw_out"])


def test_oracle_remap_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for _ in range(4):
        H, W = rng.integers(6, 60), rng.integers(6, 80)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        mx = (rng.random((23, 31)) * (W + 40) - 20).astype(np.float32)
        my = (rng.random((23, 31)) * (H + 40) - 20).astype(np.float32)
        assert np.array_equal(R.remap_cubic_wrap_u8(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_WRAP))
    v = rng.normal(size=(50, 3))
    assert all(np.array_equal(R.rodrigues(x), cv2.Rodrigues(x)[0]) for x in v)


def _mostly_equal(a, b, what, frac=2e-3):
    """Outputs that pass through float64 trigonometry on the host: a different libm / BLAS may move a sampling
    coordinate across a 1/32-pixel rounding boundary for a handful of pixels, by one grey level."""
    a, b = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
    assert a.shape == b.shape, what
    bad = np.count_nonzero(a != b)
    assert bad <= frac * a.size and (bad == 0 or np.abs(a - b).max() <= 2), f"{what}: {bad}/{a.size} differ, max {np.abs(a - b).max()}"


def test_oracle_views_match_reference_golden():
    for (t, p), ref in zip(G["e2p_views"], G["e2p_out"]):
        _mostly_equal(R.get_perspective(G["pano"], 90, t, p, 16, 16), ref, f"GetPerspective {t},{p}")
    for p, ref, mref in zip(G["p2e_phis"], G["p2e_out"], G["p2e_mask"]):
        out, mask = R.get_equirec(G["pers"], 90, 0, p, 32, 64)
        _mostly_equal(out, ref, f"GetEquirec {p}")
        _mostly_equal(mask, mref, f"GetEquirec mask {p}")


def test_oracle_process_equi_and_anchor_target_match_reference_golden():
    out = R.process_equi(G["pe_vid"], G["pe_th"].squeeze(), G["pe_ph"].squeeze(), pers_resolution=16)
    assert out.shape == G["pe_out"].shape and out.dtype == np.float32
    _mostly_equal(np.rint((out + 1) * 127.5), np.rint((G["pe_out"] + 1) * 127.5), "process_equi")
    m = R.process_equi(np.repeat(G["pe_mask_in"], 3, axis=1) * 1.0, G["pe_th"].squeeze(), G["pe_ph"].squeeze(), 16, back_norm=False)
    _mostly_equal(m, G["pe_mask_out"], "process_equi mask")
    assert tuple(R.get_maxrec_cord(G["maxrec_mask"])) == tuple(G["maxrec_out"])
    a, ap, tg, mk, rel, pit = R.get_anchor_target(torch.from_numpy(G["at_in"]), list(G["at_ph"]))
    assert np.array_equal(rel.numpy(), G["at_rel"]) and np.array_equal(pit.numpy(), G["at_pitch"])
    _mostly_equal(mk.numpy(), G["at_masks"], "anchor masks")
    _mostly_equal(np.rint((ap.numpy() + 1) * 127.5), np.rint((G["at_anchor_pers"] + 1) * 127.5), "anchor pers")
    assert np.allclose(a.numpy()[..., ::8, ::8], G["at_anchor_s8"], atol=1e-6)


def test_library_weight_table_is_opencvs():
    from imagine360_b200 import _lib
    t = np.zeros((1024, 16), np.int16)
    assert _lib.lib().i360_remap_cubic_table_i16(t.ctypes.data_as(ctypes.POINTER(ctypes.c_short))) == 0
    assert np.array_equal(t, R.cubic_table()) and (t.astype(np.int64).sum(1) == 32768).all()


def test_host_maps_and_rectangle_match_oracle():
    from imagine360_b200.host import preprocess as P
    for t, p in [(0, 0), (36.0, 26.565), (-108.0, -52.62), (180.0, 90.0), (72.0, -90.0)]:
        a, b = P.e2p_maps(90, t, p, 24, 24, 48, 96), R.e2p_maps(90, t, p, 24, 24, 48, 96)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for p in (0.0, 12.5, -33.25, 80.0):
        a, b = P.p2e_maps(90, 0, p, 24, 24, 48, 96), R.p2e_maps(90, 0, p, 24, 24, 48, 96)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    rng = np.random.default_rng(3)
    for _ in range(20):
        m = (rng.random((rng.integers(2, 30), rng.integers(2, 40))) < 0.8).astype(np.int64)
        assert tuple(P.get_maxrec_cord(m)) == tuple(R.get_maxrec_cord(m))
    assert tuple(P.get_maxrec_cord(G["maxrec_mask"])) == tuple(G["maxrec_out"])


def test_max_rectangle_against_brute_force():
    """the stack scan (oracle and product) finds a rectangle of maximal area that is all ones"""
    from imagine360_b200.host import preprocess as P
    rng = np.random.default_rng(9)
    for _ in range(25):
        h, w = int(rng.integers(1, 9)), int(rng.integers(1, 10))
        m = (rng.random((h, w)) < 0.7).astype(np.int64)
        best = 0
        for y0 in range(h):
            for x0 in range(w):
                for y1 in range(y0 + 1, h + 1):
                    for x1 in range(x0 + 1, w + 1):
                        if m[y0:y1, x0:x1].all():
                            best = max(best, (y1 - y0) * (x1 - x0))
        for fn in (R.get_maxrec_cord, P.get_maxrec_cord):
            ty, tx, rw, rh = (int(v) for v in fn(m))
            assert rw * rh == best
            if best:
                assert m[ty:ty + rh, tx:tx + rw].all()


def test_wrap_index_matches_python_modulo():
    p = np.arange(-300, 300)
    for n in (1, 2, 7, 64):
        assert np.array_equal(R._wrap(p, n), np.mod(p, n))


def test_pitch_fit_matches_sklearn():
    """pers2pano_vid's temporal smoothing of the pitch estimates (inference_dual_p2e.py:286-291)"""
    sk = pytest.importorskip("sklearn.linear_model")
    from imagine360_b200.host import preprocess as P
    rng = np.random.default_rng(4)
    for n in (2, 5, 16, 24):
        y = rng.normal(size=n) * 7 + 3
        x = np.arange(n).reshape(-1, 1)
        ref = sk.LinearRegression().fit(x, y).predict(x)
        assert np.array_equal(np.array(P.fit_pitch_linear(y)), ref)
        # and the closed form used when scikit-learn is absent agrees to rounding
        xm, ym = x.mean(), y.mean()
        coef = ((x[:, 0] - xm) * (y - ym)).sum() / ((x[:, 0] - xm) ** 2).sum()
        assert np.allclose(x[:, 0] * coef + (ym - xm * coef), ref, rtol=0, atol=1e-12)
