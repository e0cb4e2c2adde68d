"""CPU restatement (numpy) of the geometric pre-processing that feeds the denoising path -- SURVEY.md §8(f) row 1.

TEST INFRASTRUCTURE ONLY: imported by tests/, tools/make_golden_remap.py, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; the product path (imagine360_b200/host/preprocess.py + csrc/remap.cu) never imports it.

What is restated
  * ``cv2.remap(img_u8, mapx_f32, mapy_f32, INTER_CUBIC, borderMode=BORDER_WRAP)`` -- the call the reference makes in
    src/utils/pano_utils/Equirec2Perspec.py:61 and Perspec2Equirec.py:74.  OpenCV (third-party dependency, not vendored in
    /root/reference; version installed in the build container: 4.13.0) evaluates it in FIXED POINT for 8-bit images
    (modules/imgproc/src/imgwarp.cpp: RemapInvoker, initInterTab2D, remapBicubic<FixedPtCast<int,uchar,15>>):
      - the float maps are quantised to 1/32 pixel:  s = cvRound(map * 32)  (round half to even),
        integer part s >> 5, fraction index s & 31;
      - 32 x 32 tables of 4 x 4 int16 weights = round(wy * wx * 32768) with A = -0.75 cubic kernels evaluated in float32,
        the table sum forced to 32768 by correcting one entry of the rows/cols {2, 3} block;
      - out = saturate_u8((sum_16taps(src * w) + 16384) >> 15), taps wrapped with BORDER_WRAP index arithmetic.
    Pinned against the real cv2.remap on random images / maps (tests/test_remap_oracle.py, bit-exact) and through the
    committed golden vectors of tests/golden/remap_*.npz (written by tools/make_golden_remap.py from cv2 and the unmodified
    reference classes).
  * the sampling maps of ``Equirectangular.GetPerspective`` (Equirec2Perspec.py:18-62) and ``Perspective.GetEquirec``
    (Perspec2Equirec.py:29-83), float64 numpy exactly like the reference, cast to float32 at the end;
  * ``process_equi`` (inference_dual_p2e.py:113-144), the per-frame loop of ``pers2pano_vid`` (:293-301) and
    ``get_maxrec_cord`` (src/modules/utils.py:39-73).
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS


# ------------------------------------------------------------------------------------------------------
# cv2.remap, INTER_CUBIC, BORDER_WRAP, 8-bit
# ------------------------------------------------------------------------------------------------------
def _cubic_coeffs_f32(x: np.float32) -> np.ndarray:
    """interpolateCubic (imgwarp.cpp), evaluated in float32 like the C++ code."""
    f = np.float32
    A = f(-0.75)
    x = f(x)
    c = np.zeros(4, np.float32)
    c[0] = ((A * (x + f(1)) - f(5) * A) * (x + f(1)) + f(8) * A) * (x + f(1)) - f(4) * A
    c[1] = ((A + f(2)) * x - (A + f(3))) * x * x + f(1)
    c[2] = ((A + f(2)) * (f(1) - x) - (A + f(3))) * (f(1) - x) * (f(1) - x) + f(1)
    c[3] = f(1) - c[0] - c[1] - c[2]
    return c


_TAB = None


def cubic_table() -> np.ndarray:
    """[32*32, 16] int16 fixed-point weights, index = fy * 32 + fx, entry = ky * 4 + kx (initInterTab2D, fixpt)."""
    global _TAB
    if _TAB is not None:
        return _TAB
    scale = np.float32(1.0) / np.float32(INTER_TAB_SIZE)
    tab1 = np.stack([_cubic_coeffs_f32(np.float32(i) * scale) for i in range(INTER_TAB_SIZE)])   # [32, 4] float32
    out = np.zeros((INTER_TAB_SIZE * INTER_TAB_SIZE, 16), np.int16)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = (tab1[i][:, None] * tab1[j][None, :]).astype(np.float32)                    # float32 products
            it = np.clip(np.rint((v * np.float32(COEF_SCALE)).astype(np.float32)), -32768, 32767).astype(np.int32)
            isum = int(it.sum())
            if isum != COEF_SCALE:
                diff = isum - COEF_SCALE
                mk = Mk = (2, 2)
                for k1 in (2, 3):
                    for k2 in (2, 3):
                        if it[k1, k2] < it[mk]:
                            mk = (k1, k2)
                        elif it[k1, k2] > it[Mk]:
                            Mk = (k1, k2)
                if diff < 0:
                    it[Mk] -= diff
                else:
                    it[mk] -= diff
            out[i * INTER_TAB_SIZE + j] = it.reshape(16).astype(np.int16)
    _TAB = out
    return out


def _wrap(p: np.ndarray, n: int) -> np.ndarray:
    """borderInterpolate(p, n, BORDER_WRAP): C integer arithmetic (truncating division)."""
    p = p.astype(np.int64).copy()
    neg = p < 0
    q = np.trunc((p[neg] - n + 1) / n).astype(np.int64)          # C '/' truncates toward zero
    p[neg] -= q * n
    big = p >= n
    p[big] %= n
    return p


def quantise_maps(mapx: np.ndarray, mapy: np.ndarray):
    """float32 maps -> (ix, iy, fidx): integer sample position and the 32x32 fraction-table index."""
    sx = np.rint(mapx.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)     # cvRound: half to even
    sy = np.rint(mapy.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    fidx = (sy & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE + (sx & (INTER_TAB_SIZE - 1))
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    return ix, iy, fidx


def remap_cubic_wrap_u8(img: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    """cv2.remap(img, mapx, mapy, cv2.INTER_CUBIC, borderMode=cv2.BORDER_WRAP) for uint8 [H, W, C] images."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, C = img.shape
    ix, iy, fidx = quantise_maps(mapx, mapy)
    w = cubic_table()[fidx].astype(np.int64)                     # [h, w, 16]
    acc = np.zeros(mapx.shape + (C,), np.int64)
    src = img.astype(np.int64)
    for ky in range(4):
        yy = _wrap(iy - 1 + ky, H)
        for kx in range(4):
            xx = _wrap(ix - 1 + kx, W)
            acc += src[yy, xx] * w[..., ky * 4 + kx][..., None]
    out = (acc + (1 << (COEF_BITS - 1))) >> COEF_BITS            # arithmetic shift = floor
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------------
# sampling maps
# ------------------------------------------------------------------------------------------------------
def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """cv2.Rodrigues(rvec)[0] for a 3-vector (float64): R = cos(t) I + (1 - cos t) r r^T + sin(t) [r]_x."""
    r = np.asarray(rvec, np.float64).reshape(3)
    theta = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c, s = np.cos(theta), np.sin(theta)
    c1 = 1.0 - c
    x, y, z = r * (1.0 / theta)          # OpenCV multiplies by the reciprocal (bit-exact against cv2.Rodrigues)
    rrt = np.array([[x * x, x * y, x * z], [x * y, y * y, y * z], [x * z, y * z, z * z]])
    rx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return c * np.eye(3) + c1 * rrt + s * rx


def e2p_maps(fov: float, theta: float, phi: float, height: int, width: int, equ_h: int, equ_w: int):
    """(lon, lat) float32 maps of Equirectangular.GetPerspective (Equirec2Perspec.py:23-57)."""
    equ_cx = (equ_w - 1) / 2.0
    equ_cy = (equ_h - 1) / 2.0
    wFOV = fov
    hFOV = float(height) / width * wFOV
    w_len = np.tan(np.radians(wFOV / 2.0))
    h_len = np.tan(np.radians(hFOV / 2.0))
    x_map = np.ones([height, width], np.float32)
    y_map = np.tile(np.linspace(-w_len, w_len, width), [height, 1])
    z_map = -np.tile(np.linspace(-h_len, h_len, height), [width, 1]).T
    D = np.sqrt(x_map ** 2 + y_map ** 2 + z_map ** 2)
    xyz = np.stack((x_map, y_map, z_map), axis=2) / np.repeat(D[:, :, np.newaxis], 3, axis=2)
    y_axis = np.array([0.0, 1.0, 0.0], np.float32)
    z_axis = np.array([0.0, 0.0, 1.0], np.float32)
    R1 = rodrigues(z_axis * np.radians(theta))
    R2 = rodrigues(np.dot(R1, y_axis) * np.radians(-phi))
    xyz = xyz.reshape([height * width, 3]).T
    xyz = np.dot(R1, xyz)
    xyz = np.dot(R2, xyz).T
    lat = np.arcsin(xyz[:, 2])
    lon = np.arctan2(xyz[:, 1], xyz[:, 0])
    lon = lon.reshape([height, width]) / np.pi * 180
    lat = -lat.reshape([height, width]) / np.pi * 180
    lon = lon / 180 * equ_cx + equ_cx
    lat = lat / 90 * equ_cy + equ_cy
    return lon.astype(np.float32), lat.astype(np.float32)


def p2e_maps(fov: float, theta: float, phi: float, pers_h: int, pers_w: int, height: int, width: int):
    """(lon_map, lat_map float32, mask int [height, width]) of Perspective.GetEquirec (Perspec2Equirec.py:29-79)."""
    wFOV = fov
    hFOV = float(pers_h) / pers_w * fov
    w_len = np.tan(np.radians(wFOV / 2.0))
    h_len = np.tan(np.radians(hFOV / 2.0))
    x, y = np.meshgrid(np.linspace(-180, 180, width), np.linspace(90, -90, height))
    x_map = np.cos(np.radians(x)) * np.cos(np.radians(y))
    y_map = np.sin(np.radians(x)) * np.cos(np.radians(y))
    z_map = np.sin(np.radians(y))
    xyz = np.stack((x_map, y_map, z_map), axis=2)
    y_axis = np.array([0.0, 1.0, 0.0], np.float32)
    z_axis = np.array([0.0, 0.0, 1.0], np.float32)
    R1 = rodrigues(z_axis * np.radians(theta))
    R2 = rodrigues(np.dot(R1, y_axis) * np.radians(-phi))
    R1 = np.linalg.inv(R1)
    R2 = np.linalg.inv(R2)
    xyz = xyz.reshape([height * width, 3]).T
    xyz = np.dot(R2, xyz)
    xyz = np.dot(R1, xyz).T
    xyz = xyz.reshape([height, width, 3])
    inverse_mask = np.where(xyz[:, :, 0] > 0, 1, 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        xyz[:, :] = xyz[:, :] / np.repeat(xyz[:, :, 0][:, :, np.newaxis], 3, axis=2)
    inside = (-w_len < xyz[:, :, 1]) & (xyz[:, :, 1] < w_len) & (-h_len < xyz[:, :, 2]) & (xyz[:, :, 2] < h_len)
    lon_map = np.where(inside, (xyz[:, :, 1] + w_len) / 2 / w_len * pers_w, 0)
    lat_map = np.where(inside, (-xyz[:, :, 2] + h_len) / 2 / h_len * pers_h, 0)
    mask = np.where(inside, 1, 0) * inverse_mask
    return lon_map.astype(np.float32), lat_map.astype(np.float32), mask


# ------------------------------------------------------------------------------------------------------
# the reference functions built on them
# ------------------------------------------------------------------------------------------------------
def get_perspective(pano_u8: np.ndarray, fov, theta, phi, height, width) -> np.ndarray:
    lon, lat = e2p_maps(fov, theta, phi, height, width, pano_u8.shape[0], pano_u8.shape[1])
    return remap_cubic_wrap_u8(pano_u8, lon, lat)


def get_equirec(pers_u8: np.ndarray, fov, theta, phi, height, width):
    """-> (persp * mask [height, width, 3] (int64 like the reference's uint8 * int product), mask [height, width, 3])."""
    lon, lat, mask = p2e_maps(fov, theta, phi, pers_u8.shape[0], pers_u8.shape[1], height, width)
    persp = remap_cubic_wrap_u8(pers_u8, lon, lat)
    mask3 = np.repeat(mask[:, :, np.newaxis], 3, axis=2)
    return persp * mask3, mask3


def process_equi(panovid: np.ndarray, thetas, phis, pers_resolution=256, back_norm=True) -> np.ndarray:
    """inference_dual_p2e.py:113-144 on a float32 array [f, c, h, w]; returns float32 [f, m, c, h, w]."""
    x = (panovid + 1) * 127.5 if back_norm else panovid * 255
    out = []
    for i in range(x.shape[0]):
        pano = np.transpose(x[i], (1, 2, 0)).astype(np.uint8)
        views = []
        for th, ph in zip(thetas, phis):
            img = get_perspective(pano, 90, th, ph, pers_resolution, pers_resolution)
            views.append((img.astype(np.float32) / 127.5) - 1 if back_norm
                         else np.expand_dims(np.any(img > 0, axis=-1), axis=-1))
        out.append(np.stack(views))
    out = np.stack(out, axis=0).astype(np.float32)
    return np.transpose(out, (0, 1, 4, 2, 3))


def pers2pano_frames(persframes: np.ndarray, ph_list, pano_h=256, pano_w=512, fov=90, th=0):
    """per-frame loop of pers2pano_vid (inference_dual_p2e.py:293-301): -> (frames uint8 [f,H,W,3], mask uint8 [f,H,W,1])
    where mask = 1 outside the projected view."""
    frames, masks = [], []
    for i in range(persframes.shape[0]):
        pano, mask = get_equirec(persframes[i], fov, th, ph_list[i], pano_h, pano_w)
        frames.append(pano.astype(np.uint8))
        m = np.any((1 - mask) > 0, axis=-1).astype(np.uint8)
        masks.append(m[..., None])
    return np.stack(frames, axis=0), np.stack(masks, axis=0)


def get_maxrec_cord(mask: np.ndarray):
    """Largest all-ones axis-aligned rectangle of a 0/1 mask -> (top_left_y, top_left_x, rect_width, rect_height).

    Same algorithm and candidate order as src/modules/utils.py:39-73 (column run-lengths, then per row a monotone stack
    over the run-length histogram, strict '>' when comparing areas), so that ties pick the same rectangle."""
    ones = np.asarray(mask) == 1
    n_rows, n_cols = ones.shape
    runs = np.zeros((n_rows, n_cols), dtype=np.int64)            # consecutive ones ending at (row, col), going up
    for r in range(n_rows):
        runs[r] = np.where(ones[r], (runs[r - 1] if r else 0) + 1, 0)
    best_area, best = 0, (0, 0, 0, 0)
    for r in range(n_rows):
        hist = runs[r]
        open_cols: list[int] = []                                 # columns with increasing bar heights
        for c in range(n_cols + 1):
            cur = hist[c] if c < n_cols else 0
            while open_cols and cur < hist[open_cols[-1]]:
                bar = int(hist[open_cols.pop()])
                left = open_cols[-1] + 1 if open_cols else 0
                span = c - left
                if bar * span > best_area:
                    best_area, best = bar * span, (r - bar + 1, left, span, bar)
            open_cols.append(c)
    return best


def get_anchor_target(pixel_values, ph_list, fov=90, th=0):
    """animatediff/utils/video_mask.py:158-217 on torch tensors (CPU), built from the numpy restatements above.
    pixel_values [f, 3, h, w] or [b, f, 3, h, w] in (-1, 1)."""
    import torch
    import torch.nn.functional as F

    if pixel_values.dim() == 4:
        pixel_values = pixel_values.unsqueeze(0)
    b, f, c, h, w = pixel_values.shape
    size = int(h / 2)
    crops = []
    for i in range(f):
        frame = (pixel_values[0, i].permute(1, 2, 0).cpu().numpy() + 1) / 2 * 255
        crops.append(get_perspective(frame.astype(np.uint8), fov, th, ph_list[i], size, size).astype(np.uint8))
    crops = np.stack(crops)
    anchor_pers = torch.from_numpy((crops / 127.5) - 1).permute(0, 3, 1, 2).unsqueeze(0).expand(b, -1, -1, -1, -1)
    masks, anchors, rel, pitchs = [], [], [], []
    for i in range(f):
        _, _, inside = p2e_maps(fov, th, ph_list[i], size, size, h, w)
        masks.append(torch.from_numpy(1 - inside)[None, None].expand(b, -1, -1, -1).float())
        ty, tx, rw, rh = get_maxrec_cord(inside)
        crop = pixel_values[:, i, :, ty:ty + rh, tx:tx + rw]
        anchors.append(F.interpolate(crop, size=(256, 256), mode="bilinear", align_corners=False))
        pitchs.append(torch.tensor([ph_list[i]]))
        rel.append(torch.tensor([int(h / 2 - (ty + ty + rh) / 2), int(w / 2 - (tx + tx + rw) / 2), rh, rw, h, w]))
    return (torch.stack(anchors, dim=1), anchor_pers, pixel_values.clone(), torch.stack(masks, dim=1),
            torch.stack(rel, dim=0).unsqueeze(0).repeat(b, 1, 1), torch.stack(pitchs, dim=1))


def video_to_frames_u8(videos: np.ndarray, rescale: bool = False) -> np.ndarray:
    """uint8 conversion of save_videos_grid (animatediff/utils/util.py:55-72) for one video [1, 3, t, h, w] float32 ->
    [t, h, w, 3] (make_grid of a single image is the identity)."""
    x = np.transpose(videos[0].astype(np.float32), (1, 2, 3, 0))
    if rescale:
        x = (x + np.float32(1.0)) / np.float32(2.0)
    return (x * np.float32(255)).astype(np.uint8)
