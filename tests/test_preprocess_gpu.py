"""GPU parity of the geometric pre-processing row: integer work, so the bar is BIT-EXACT
against the numpy oracle (itself pinned to cv2.remap and the unmodified reference functions, test_remap_oracle.py)
and against the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "remap_golden.npz"))


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_remap_kernel_bit_exact_against_cv2_golden():
    from imagine360_b200.host import preprocess as P
    out, _ = P.remap_cubic_wrap(_cuda(G["raw_img"])[None], _cuda(G["raw_mx"])[None], _cuda(G["raw_my"])[None])
    assert np.array_equal(out[0, 0].cpu().numpy(), G["raw_out"])


@pytest.mark.parametrize("n,H,W,m,h,w,paired", [(3, 37, 53, 4, 20, 33, False), (5, 16, 16, 5, 40, 70, True), (2, 128, 256, 20, 64, 64, False),
                                                (1, 5, 7, 1, 9, 9, False)])
def test_remap_kernel_bit_exact_against_oracle(n, H, W, m, h, w, paired):
    """random frames, maps that run far outside the image (BORDER_WRAP on both axes), keep-mask, both output forms"""
    from imagine360_b200.host import preprocess as P
    from oracle import remap as R
    rng = np.random.default_rng(n * 1000 + H)
    src = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    mx = (rng.random((m, h, w)) * 3 * W - W).astype(np.float32)
    my = (rng.random((m, h, w)) * 3 * H - H).astype(np.float32)
    keep = (rng.random((m, h, w)) < 0.7).astype(np.uint8)
    u8, f32 = P.remap_cubic_wrap(_cuda(src), _cuda(mx), _cuda(my), keep=_cuda(keep), paired=paired, f32_mode=1)
    u8, f32 = u8.cpu().numpy(), f32.cpu().numpy()
    for i in range(n):
        for v in ([i] if paired else range(m)):
            ref = R.remap_cubic_wrap_u8(src[i], mx[v], my[v]) * keep[v][..., None]
            got = u8[i, 0 if paired else v]
            assert np.array_equal(got, ref), (i, v, np.count_nonzero(got != ref))
            assert np.array_equal(f32[i, 0 if paired else v], np.transpose((ref.astype(np.float32) / 127.5) - 1, (2, 0, 1)))
    _, anym = P.remap_cubic_wrap(_cuda(src), _cuda(mx), _cuda(my), keep=_cuda(keep), paired=paired, want_u8=False, f32_mode=2)
    assert np.array_equal(anym.cpu().numpy()[:, :, 0], (u8 > 0).any(-1).astype(np.float32))


def test_process_equi_matches_oracle_and_reference_golden():
    from imagine360_b200.host import preprocess as P
    from oracle import remap as R
    rng = np.random.default_rng(11)
    vid = (rng.random((4, 3, 64, 128)).astype(np.float32) * 2 - 1)
    th = np.array([[0.0, 72.0, -144.0, 36.0, 180.0]]); ph = np.array([[0.0, 26.565, -52.62, 90.0, -26.565]])
    out = P.process_equi(torch.from_numpy(vid), th, ph, pers_resolution=32)           # CPU tensor in -> CPU tensor out
    assert out.device.type == "cpu" and out.dtype == torch.float32
    assert np.array_equal(out.numpy(), R.process_equi(vid, th.squeeze(), ph.squeeze(), 32))
    msk = (rng.random((4, 1, 64, 128)) < 0.5).astype(np.uint8)
    out = P.process_equi(torch.from_numpy(msk).cuda().repeat(1, 3, 1, 1), th, ph, pers_resolution=32, back_norm=False)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), R.process_equi(np.repeat(msk, 3, 1) * 1.0, th.squeeze(), ph.squeeze(), 32, False))
    # the unmodified reference function, recorded in the build container (host trigonometry may differ by an ulp
    # between machines: allow a handful of one-grey-level differences)
    got = P.process_equi(torch.from_numpy(G["pe_vid"]), G["pe_th"], G["pe_ph"], pers_resolution=16).numpy()
    a, b = np.rint((got + 1) * 127.5), np.rint((G["pe_out"] + 1) * 127.5)
    assert np.count_nonzero(a != b) <= 2e-3 * a.size and np.abs(a - b).max() <= 2
    gm = P.process_equi(torch.from_numpy(G["pe_mask_in"]).repeat(1, 3, 1, 1), G["pe_th"], G["pe_ph"], pers_resolution=16, back_norm=False)
    assert np.count_nonzero(gm.numpy() != G["pe_mask_out"]) <= 2e-3 * gm.numel()


def test_pers2pano_frames_matches_oracle():
    from imagine360_b200.host import preprocess as P
    from oracle import remap as R
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, (3, 48, 48, 3), dtype=np.uint8)
    phs = [0.0, 9.75, -21.5]
    pano, mask = P.pers2pano_frames(frames, phs, pano_H=64, pano_W=128)
    rp, rm = R.pers2pano_frames(frames, phs, 64, 128)
    assert isinstance(pano, np.ndarray) and pano.dtype == np.uint8 and mask.shape == (3, 64, 128, 1)
    assert np.array_equal(pano, rp) and np.array_equal(mask, rm)


def test_get_anchor_target_matches_oracle_and_reference_golden():
    from imagine360_b200.host import preprocess as P
    from oracle import remap as R
    pv = torch.from_numpy(G["at_in"])
    phl = [float(x) for x in G["at_ph"]]
    got = P.get_anchor_target(pv.cuda(), phl)
    ref = R.get_anchor_target(pv, phl)
    names = ["anchor", "anchor_pers", "target", "masks", "rel", "pitch"]
    for n, a, b in zip(names, got, ref):
        assert a.shape == b.shape and a.dtype == b.dtype, (n, a.shape, b.shape, a.dtype, b.dtype)
        if n == "anchor":                                 # F.interpolate on the GPU vs on the CPU: float rounding only
            assert torch.allclose(a.cpu(), b, atol=1e-5)
        else:
            assert torch.equal(a.cpu(), b), n
    assert np.array_equal(got[4].cpu().numpy(), G["at_rel"])
    assert np.allclose(got[0].cpu().numpy()[..., ::8, ::8], G["at_anchor_s8"], atol=1e-5)


def test_reference_named_classes_match_golden():
    """Equirectangular.GetPerspective / Perspective.GetEquirec mirrors against outputs of the unmodified classes."""
    from imagine360_b200.host import preprocess as P
    from oracle import remap as R
    eq = P.Equirectangular(G["pano"])
    for (t, p), ref in zip(G["e2p_views"], G["e2p_out"]):
        got = eq.GetPerspective(90, t, p, 16, 16)
        assert got.dtype == np.uint8 and np.array_equal(got, R.get_perspective(G["pano"], 90, t, p, 16, 16))
        assert np.count_nonzero(got != ref) <= 2
    for p, ref, mref in zip(G["p2e_phis"], G["p2e_out"], G["p2e_mask"]):
        got, mask = P.Perspective(G["pers"], 90, 0, p).GetEquirec(32, 64)
        assert got.dtype == ref.dtype and mask.dtype == mref.dtype and got.shape == ref.shape
        assert np.count_nonzero(got != ref) <= 4 and np.count_nonzero(mask != mref) <= 3


@pytest.mark.parametrize("rescale", [False, True])
def test_video_to_frames_u8_bit_exact(rescale):
    """output side (8(f) row 3): save_videos_grid's uint8 conversion, against the oracle and against torch's own ops"""
    from imagine360_b200.host import preprocess as P
    from oracle import remap as R
    g = torch.Generator().manual_seed(5)
    v = torch.rand(1, 3, 5, 24, 40, generator=g)
    if rescale:
        v = v * 2 - 1
    v[0, 0, 0, 0, :4] = torch.tensor([1.0, 0.0, 0.999999, 0.5]) if not rescale else torch.tensor([1.0, -1.0, 0.999999, 0.0])
    out = P.video_to_frames_u8(v.cuda(), rescale)
    assert out.dtype == torch.uint8 and out.shape == (5, 24, 40, 3)
    assert np.array_equal(out.cpu().numpy(), R.video_to_frames_u8(v.numpy(), rescale))
    x = v[0].permute(1, 0, 2, 3)
    x = (x + 1.0) / 2.0 if rescale else x
    ref = (x.permute(0, 2, 3, 1) * 255).numpy().astype(np.uint8)        # the reference's expression on one frame at a time
    assert np.array_equal(out.cpu().numpy(), ref)


def test_pers2pano_vid_with_a_stub_camera_model():
    """reference signature: per-frame pitch from the caller's GeoCalib-like model, least-squares fit over the frame
    index, then one batched perspective -> panorama resampling; against the oracle fed with the same fitted pitches"""
    import types
    from imagine360_b200.host import preprocess as P
    from oracle import remap as R
    rng = np.random.default_rng(21)
    frames = rng.integers(0, 256, (4, 32, 32, 3), dtype=np.uint8)
    pitches_deg = [3.0, 4.5, 2.0, 6.0]

    class Stub:
        def __init__(self):
            self.i = 0

        def calibrate(self, img):
            assert img.is_cuda and img.shape == (3, 32, 32) and float(img.max()) <= 1.0
            rp = torch.deg2rad(torch.tensor([0.0, pitches_deg[self.i]], dtype=torch.float64))
            self.i += 1
            return {"camera": None, "gravity": types.SimpleNamespace(rp=rp)}

    ph_list, pano, mask, pano2 = P.pers2pano_vid(Stub(), "geocalib", frames, pano_H=48, pano_W=96)
    assert len(ph_list) == 4 and pano is pano2 and pano.shape == (4, 48, 96, 3) and mask.shape == (4, 48, 96, 1)
    x = np.arange(4)
    slope, icpt = np.polyfit(x, np.rad2deg(np.deg2rad(np.array(pitches_deg))), 1)
    assert np.allclose(ph_list, slope * x + icpt, atol=1e-9)
    rp, rm = R.pers2pano_frames(frames, ph_list, 48, 96)
    assert np.array_equal(pano, rp) and np.array_equal(mask, rm)
    # no model: constant pitch for every frame
    ph2, pano3, _, _ = P.pers2pano_vid(None, "none", frames[:2], pano_H=48, pano_W=96, ph=5.0)
    assert ph2 == [5.0, 5.0] and np.array_equal(pano3, R.pers2pano_frames(frames[:2], ph2, 48, 96)[0])
