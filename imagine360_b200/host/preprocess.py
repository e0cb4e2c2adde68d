"""GPU pre-processing of the geometric warps that feed the denoising path.

Host-side mirror of the reference functions that resample 8-bit frames between the equirectangular panorama and
perspective views with ``cv2.remap(INTER_CUBIC, BORDER_WRAP)`` on the CPU, one (frame, view) pair at a time:

  * ``process_equi``        inference_dual_p2e.py:113-144   (F x 20 remaps, maps recomputed for every pair)
  * ``pers2pano_vid`` / ``pers2pano_frames``  inference_dual_p2e.py:256-305 (the camera estimation models it calls are
                                                              third-party objects passed in by the caller, as in the reference)
  * ``get_anchor_target``   animatediff/utils/video_mask.py:158-217
  * ``get_maxrec_cord``     src/modules/utils.py:39-73      (pure-Python O(H*W) scan -> numpy run-lengths + stack)

Same names, argument meaning and return layouts.  The sampling maps depend only on (FoV, theta, phi, sizes): they are
built once on the host in float64 with the reference's own operation sequence (``Equirec2Perspec.py:23-57``,
``Perspec2Equirec.py:29-72``; cv2.Rodrigues restated in numpy, bit-exact), cached, and uploaded; every frame of every
view is then resampled by ONE launch of ``i360_remap_cubic_wrap_u8``, which reproduces OpenCV's fixed-point bicubic
arithmetic bit for bit.  There is no CPU fallback: without the CUDA library these functions raise.
"""
from __future__ import annotations

import ctypes
from ctypes import c_int, c_longlong, c_void_p

import numpy as np
import torch
import torch.nn.functional as F

from .._lib import check, lib

_maps: dict = {}
_tables: dict = {}


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_tensor_device(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current (kernels launch on the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        for t in a:
            if torch.is_tensor(t) and t.is_cuda:
                with torch.cuda.device(t.device):
                    return fn(*a, **k)
        return fn(*a, **k)
    return wrapped


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def cubic_table(device) -> torch.Tensor:
    """OpenCV's fixed-point bicubic weights [1024, 16] int16 on ``device`` (built by the library, cached)."""
    key = str(device)
    if key not in _tables:
        host = np.zeros((1024, 16), np.int16)
        check(lib().i360_remap_cubic_table_i16(host.ctypes.data_as(ctypes.POINTER(ctypes.c_short))), "i360_remap_cubic_table_i16")
        _tables[key] = torch.from_numpy(host).to(device)
    return _tables[key]


# ------------------------------------------------------------------------------------------------------
# sampling maps (host, float64, cached)
# ------------------------------------------------------------------------------------------------------
def _rodrigues(rvec) -> np.ndarray:
    """Rotation matrix of a rotation vector, the formula and operation order of cv2.Rodrigues (calib3d)."""
    r = np.asarray(rvec, np.float64).reshape(3)
    angle = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if angle < np.finfo(np.float64).eps:
        return np.eye(3)
    co, si = np.cos(angle), np.sin(angle)
    ax, ay, az = r * (1.0 / angle)
    outer = np.array([[ax * ax, ax * ay, ax * az], [ax * ay, ay * ay, ay * az], [ax * az, ay * az, az * az]])
    cross = np.array([[0, -az, ay], [az, 0, -ax], [-ay, ax, 0]])
    return co * np.eye(3) + (1.0 - co) * outer + si * cross


def _view_rotations(theta, phi):
    up = np.array([0.0, 0.0, 1.0], np.float32)
    side = np.array([0.0, 1.0, 0.0], np.float32)
    yaw = _rodrigues(up * np.radians(theta))
    pitch = _rodrigues(np.dot(yaw, side) * np.radians(-phi))
    return yaw, pitch


def e2p_maps(fov, theta, phi, height: int, width: int, equ_h: int, equ_w: int):
    """Pixel coordinates (x, y float32 [height, width]) in the panorama for every pixel of a perspective view
    (Equirec2Perspec.py:23-57)."""
    cx, cy = (equ_w - 1) / 2.0, (equ_h - 1) / 2.0
    half_w = np.tan(np.radians(fov / 2.0))
    half_h = np.tan(np.radians(float(height) / width * fov / 2.0))
    fwd = np.ones([height, width], np.float32)
    right = np.tile(np.linspace(-half_w, half_w, width), [height, 1])
    upw = -np.tile(np.linspace(-half_h, half_h, height), [width, 1]).T
    length = np.sqrt(fwd ** 2 + right ** 2 + upw ** 2)
    rays = np.stack((fwd, right, upw), axis=2) / np.repeat(length[:, :, np.newaxis], 3, axis=2)
    yaw, pitch = _view_rotations(theta, phi)
    rays = rays.reshape([height * width, 3]).T
    rays = np.dot(pitch, np.dot(yaw, rays)).T
    lat = -np.arcsin(rays[:, 2]).reshape([height, width]) / np.pi * 180
    lon = np.arctan2(rays[:, 1], rays[:, 0]).reshape([height, width]) / np.pi * 180
    return (lon / 180 * cx + cx).astype(np.float32), (lat / 90 * cy + cy).astype(np.float32)


def p2e_maps(fov, theta, phi, pers_h: int, pers_w: int, height: int, width: int):
    """(x, y float32, inside uint8) [height, width]: where every panorama pixel samples the perspective image, and
    whether it is covered by the view at all (Perspec2Equirec.py:29-72, :76)."""
    half_w = np.tan(np.radians(fov / 2.0))
    half_h = np.tan(np.radians(float(pers_h) / pers_w * fov / 2.0))
    lon, lat = np.meshgrid(np.linspace(-180, 180, width), np.linspace(90, -90, height))
    dirs = np.stack((np.cos(np.radians(lon)) * np.cos(np.radians(lat)), np.sin(np.radians(lon)) * np.cos(np.radians(lat)),
                     np.sin(np.radians(lat))), axis=2)
    yaw, pitch = _view_rotations(theta, phi)
    yaw_inv, pitch_inv = np.linalg.inv(yaw), np.linalg.inv(pitch)
    dirs = dirs.reshape([height * width, 3]).T
    dirs = np.dot(yaw_inv, np.dot(pitch_inv, dirs)).T.reshape([height, width, 3])
    front = np.where(dirs[:, :, 0] > 0, 1, 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        dirs[:, :] = dirs[:, :] / np.repeat(dirs[:, :, 0][:, :, np.newaxis], 3, axis=2)
    hit = (-half_w < dirs[:, :, 1]) & (dirs[:, :, 1] < half_w) & (-half_h < dirs[:, :, 2]) & (dirs[:, :, 2] < half_h)
    x = np.where(hit, (dirs[:, :, 1] + half_w) / 2 / half_w * pers_w, 0)
    y = np.where(hit, (-dirs[:, :, 2] + half_h) / 2 / half_h * pers_h, 0)
    return x.astype(np.float32), y.astype(np.float32), (np.where(hit, 1, 0) * front).astype(np.uint8)


def _cached_e2p(fov, thetas, phis, res, H, W, device):
    key = ("e2p", float(fov), tuple(float(t) for t in thetas), tuple(float(p) for p in phis), res, H, W, str(device))
    if key not in _maps:
        xy = [e2p_maps(fov, t, p, res, res, H, W) for t, p in zip(thetas, phis)]
        _maps[key] = (torch.from_numpy(np.stack([m[0] for m in xy])).to(device), torch.from_numpy(np.stack([m[1] for m in xy])).to(device))
        if len(_maps) > 64:
            _maps.pop(next(iter(_maps)))
    return _maps[key]


def _cached_p2e(fov, theta, phis, ph, pw, H, W, device):
    key = ("p2e", float(fov), float(theta), tuple(float(p) for p in phis), ph, pw, H, W, str(device))
    if key not in _maps:
        m = [p2e_maps(fov, theta, p, ph, pw, H, W) for p in phis]
        _maps[key] = tuple(torch.from_numpy(np.stack([q[i] for q in m])).to(device) for i in range(3))
        if len(_maps) > 64:
            _maps.pop(next(iter(_maps)))
    return _maps[key]


# ------------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------------
@_on_tensor_device
def frames_to_u8(x: torch.Tensor, back_norm: bool) -> torch.Tensor:
    """[n, 3, H, W] -> uint8 [n, H, W, 3] like ``((x + 1) * 127.5 | x * 255)`` followed by ``.astype(np.uint8)``."""
    n, c, H, W = x.shape
    assert c == 3
    if x.dtype != torch.float32:      # integer masks (uint8 0/1 * 255 in the reference): exact in any arithmetic
        y = (x + 1) * 127.5 if back_norm else x * 255
        return y.permute(0, 2, 3, 1).to(torch.uint8).contiguous()
    x = x.contiguous()
    out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=x.device)
    check(lib().i360_frames_to_u8_nhwc(_p(x), c_longlong(3 * H * W), c_longlong(H * W), _p(out), c_longlong(n), c_int(H), c_int(W),
                                       c_int(1 if back_norm else 0), _stream()), "i360_frames_to_u8_nhwc")
    return out


@_on_tensor_device
def video_to_frames_u8(videos: torch.Tensor, rescale: bool = False) -> torch.Tensor:
    """Output side: the uint8 conversion of ``save_videos_grid`` (animatediff/utils/util.py:55-72)
    on the GPU for a single video: videos [1, c=3, t, h, w] float32 in (0, 1) (or (-1, 1) with ``rescale``) ->
    uint8 [t, h, w, 3], bit-identical to ``((x + 1) / 2 if rescale else x) * 255`` + ``.numpy().astype(np.uint8)``.
    (``make_grid`` is the identity for one video; the mp4 container is written by imageio on the CPU.)"""
    b, c, t, h, w = videos.shape
    if b != 1 or c != 3:
        raise NotImplementedError("grid layout of several videos is CPU post-processing (out of scope)")
    dev = _device_of(videos)
    x = videos.to(dev, torch.float32).contiguous()
    out = torch.empty((t, h, w, 3), dtype=torch.uint8, device=dev)
    check(lib().i360_frames_to_u8_nhwc(_p(x), c_longlong(h * w), c_longlong(t * h * w), _p(out), c_longlong(t), c_int(h), c_int(w),
                                       c_int(2 if rescale else 0), _stream()), "i360_frames_to_u8_nhwc")
    return out


def save_videos_grid(videos: torch.Tensor, path: str, rescale=False, n_rows=6, fps=8, n_frames=None):
    """Reference signature; frames are converted on the GPU, one D2H copy of uint8 instead of float32, then imageio."""
    import os
    pre = getattr(videos, "_i360_frames_u8", None)       # AnimationPipeline output: frames already converted on the GPU
    if pre is not None and not rescale and tuple(pre.shape) == (videos.shape[2], videos.shape[3], videos.shape[4], 3):
        frames = pre
    else:
        frames = video_to_frames_u8(videos, rescale)
    if n_frames is not None:
        frames = frames[:n_frames]
    frames = list(frames.cpu().numpy())
    os.makedirs(os.path.dirname(path), exist_ok=True)
    import imageio
    imageio.mimsave(path, frames, fps=fps)


@_on_tensor_device
def remap_cubic_wrap(src_u8: torch.Tensor, mapx: torch.Tensor, mapy: torch.Tensor, keep: torch.Tensor | None = None,
                     paired: bool = False, want_u8: bool = True, f32_mode: int = 0):
    """Batched ``cv2.remap(INTER_CUBIC, BORDER_WRAP)``: src uint8 [n, H, W, 3]; maps float32 [m, h, w].
    -> (uint8 [n, m|1, h, w, 3] or None, float32 [n, m|1, 3|1, h, w] or None)."""
    assert src_u8.dtype == torch.uint8 and src_u8.is_cuda and src_u8.is_contiguous() and src_u8.shape[-1] == 3
    assert mapx.dtype == torch.float32 and mapx.is_contiguous() and mapy.is_contiguous() and mapx.shape == mapy.shape
    n, H, W, _ = src_u8.shape
    m, h, w = mapx.shape
    n_out = 1 if paired else m
    dev = src_u8.device
    out_u8 = torch.empty((n, n_out, h, w, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    out_f = torch.empty((n, n_out, 3 if f32_mode == 1 else 1, h, w), dtype=torch.float32, device=dev) if f32_mode else None
    if keep is not None:
        assert keep.dtype == torch.uint8 and keep.shape == mapx.shape and keep.is_contiguous()
    rc = lib().i360_remap_cubic_wrap_u8(_p(src_u8), c_int(n), c_int(H), c_int(W), _p(mapx), _p(mapy), _p(keep), c_int(m), c_int(h),
                                        c_int(w), c_int(1 if paired else 0), _p(cubic_table(dev)), _p(out_u8), _p(out_f),
                                        c_int(f32_mode), _stream())
    check(rc, "i360_remap_cubic_wrap_u8")
    return out_u8, out_f


# ------------------------------------------------------------------------------------------------------
# reference-named entry points
# ------------------------------------------------------------------------------------------------------
def _device_of(x: torch.Tensor):
    return x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())


def process_equi(panovid_data, thetas, phis, pers_resolution=256, back_norm=True):
    """panovid_data [f, c, h, w] in (-1, 1) (back_norm) or (0, 1) -> persvid_data [f, m, c, h, w] float32 (m views;
    c = 1 visibility mask when ``back_norm`` is False).  Returned on the device of the input."""
    thetas = np.asarray(thetas).squeeze().reshape(-1)
    phis = np.asarray(phis).squeeze().reshape(-1)
    dev = _device_of(panovid_data)
    f, c, H, W = panovid_data.shape
    frames = frames_to_u8(panovid_data.to(dev), back_norm)
    mx, my = _cached_e2p(90, thetas, phis, pers_resolution, H, W, dev)
    _, out = remap_cubic_wrap(frames, mx, my, want_u8=False, f32_mode=1 if back_norm else 2)
    return out.to(panovid_data.device)


def pers2pano_frames(persframes, ph_list, pano_H=256, pano_W=512, fov=90, th=0):
    """Per-frame loop of ``pers2pano_vid``: persframes uint8 [f, h, w, 3] (numpy or tensor), one pitch per frame ->
    (pano_frames uint8 [f, H, W, 3], pano_mask uint8 [f, H, W, 1]; mask = 1 where the view does NOT cover the panorama),
    numpy arrays like the reference."""
    as_np = not torch.is_tensor(persframes)
    src = torch.from_numpy(np.ascontiguousarray(persframes)) if as_np else persframes
    dev = _device_of(src)
    src = src.to(dev).contiguous()
    f, ph, pw, _ = src.shape
    mx, my, inside = _cached_p2e(fov, th, [float(p) for p in ph_list], ph, pw, pano_H, pano_W, dev)
    pano, _ = remap_cubic_wrap(src, mx, my, keep=inside, paired=True)
    pano = pano[:, 0]
    mask = (1 - inside).unsqueeze(-1)
    return (pano.cpu().numpy(), mask.cpu().numpy()) if as_np else (pano, mask)


def fit_pitch_linear(pitches) -> list:
    """The temporal smoothing of the per-frame pitch estimates in ``pers2pano_vid`` (inference_dual_p2e.py:286-291):
    ordinary least squares of pitch over the frame index, evaluated at the frame indices.  Uses scikit-learn's
    ``LinearRegression`` when it is installed (the reference's own dependency: same float64 result bit for bit), else
    the centred closed form sklearn solves."""
    y = np.asarray(pitches, dtype=np.float64).reshape(-1)
    x = np.arange(len(y), dtype=np.float64).reshape(-1, 1)
    try:
        from sklearn.linear_model import LinearRegression
        return LinearRegression().fit(x, y).predict(x).tolist()
    except ImportError:
        xm, ym = x.mean(), y.mean()
        denom = ((x[:, 0] - xm) ** 2).sum()
        coef = ((x[:, 0] - xm) * (y - ym)).sum() / denom if denom > 0 else 0.0
        return (x[:, 0] * coef + (ym - xm * coef)).tolist()


def pers2pano_vid(model, modelname, persframes, pano_H=256, pano_W=512, fov=90, th=0, ph=0):
    """Reference signature (inference_dual_p2e.py:256-305).  The camera-estimation model is the caller's (third party:
    GeoCalib / PerspectiveFields, called exactly like the reference does); the pitch fit runs on the host and the F
    perspective -> panorama resamplings are one GPU launch.  Returns (ph_list, pano_frames, pano_mask, pano_frames).
    Without a model the constant ``ph`` is used for every frame (the reference's own else-branch cannot run: it fits
    a regression on an empty index list)."""
    n = persframes.shape[0]
    estimates = []
    for i in range(n):
        if modelname == "geocalib":
            img = torch.Tensor(persframes[i] / 255).permute(2, 0, 1).cuda()
            results = model.calibrate(img)
            _roll, pitch = torch.rad2deg(results["gravity"].rp).unbind(-1)
            estimates.append(pitch.item())
        elif modelname == "perspectivefields":
            estimates.append(model.inference(img_bgr=persframes[i])["pred_pitch"].cpu().numpy())
    ph_list = fit_pitch_linear(estimates) if len(estimates) == n and n > 0 else [ph] * n
    frames, mask = pers2pano_frames(persframes, ph_list, pano_H=pano_H, pano_W=pano_W, fov=fov, th=th)
    return ph_list, frames, mask, frames


class Equirectangular:
    """``src.utils.pano_utils.Equirec2Perspec.Equirectangular`` (Equirec2Perspec.py:6-62) over the GPU remap: keeps the
    uint8 panorama on the device, ``GetPerspective`` returns a numpy uint8 view like the reference."""

    def __init__(self, img_name, text2light=False):
        if isinstance(img_name, str):
            raise NotImplementedError("pass the decoded uint8 frame (file reading is out of scope of the GPU path)")
        img = np.roll(img_name, -60, axis=0) if text2light else img_name
        self._height, self._width, _ = img.shape
        self._img = torch.from_numpy(np.ascontiguousarray(img)).cuda()[None]

    def GetPerspective(self, FOV, THETA, PHI, height, width):
        key = ("e2p1", float(FOV), float(THETA), float(PHI), height, width, self._height, self._width)
        if key not in _maps:
            mx, my = e2p_maps(FOV, THETA, PHI, height, width, self._height, self._width)
            _maps[key] = (torch.from_numpy(mx)[None].cuda(), torch.from_numpy(my)[None].cuda())
            if len(_maps) > 64:
                _maps.pop(next(iter(_maps)))
        out, _ = remap_cubic_wrap(self._img, *_maps[key])
        return out[0, 0].cpu().numpy()


class Perspective:
    """``src.utils.pano_utils.Perspec2Equirec.Perspective`` (Perspec2Equirec.py:6-83): ``GetEquirec`` -> (persp * mask
    int64 [H, W, 3], mask int64 [H, W, 3]) like the reference."""

    def __init__(self, img_name, FOV, THETA, PHI):
        if isinstance(img_name, str):
            raise NotImplementedError("pass the decoded uint8 frame (file reading is out of scope of the GPU path)")
        self._height, self._width, _ = img_name.shape
        self._img = torch.from_numpy(np.ascontiguousarray(img_name)).cuda()[None]
        self.wFOV, self.THETA, self.PHI = FOV, THETA, PHI

    def GetEquirec(self, height, width):
        mx, my, inside = _cached_p2e(self.wFOV, self.THETA, [float(self.PHI)], self._height, self._width, height, width, self._img.device)
        out, _ = remap_cubic_wrap(self._img, mx, my, keep=inside)
        mask = np.repeat(inside[0].cpu().numpy().astype(np.int64)[:, :, np.newaxis], 3, axis=2)
        return out[0, 0].cpu().numpy().astype(np.int64), mask


def get_maxrec_cord(input):
    """Largest all-ones axis-aligned rectangle of a 0/1 mask -> (top_left_y, top_left_x, rect_width, rect_height);
    candidates are visited in the reference's order so ties resolve identically."""
    if isinstance(input, torch.Tensor):
        input = input.cpu().numpy()
    ones = np.asarray(input) == 1
    rows, cols = ones.shape
    runs = np.zeros((rows, cols), dtype=np.int64)
    for r in range(rows):
        runs[r] = np.where(ones[r], (runs[r - 1] if r else 0) + 1, 0)
    best_area, best = 0, (0, 0, 0, 0)
    for r in range(rows):
        hist = runs[r].tolist() + [0]
        stack: list[int] = []
        for c in range(cols + 1):
            cur = hist[c]
            while stack and cur < hist[stack[-1]]:
                bar = hist[stack.pop()]
                left = stack[-1] + 1 if stack else 0
                if bar * (c - left) > best_area:
                    best_area, best = bar * (c - left), (r - bar + 1, left, c - left, bar)
            stack.append(c)
    return best


def get_anchor_target(pixel_values, ph_list, fov=90, th=0):
    """pixel_values [f, 3, h, w] (or [b, f, 3, h, w]) in (-1, 1), one pitch per frame ->
    (anchor_pixels_values [b, f, 3, 256, 256], anchor_pixels_values_pers [b, f, 3, h/2, h/2], target_pixels_values,
     masks [b, f, 1, h, w], relative_positions [b, f, 6], pitchs [1, f])  -- video_mask.py:158-217."""
    if len(pixel_values.shape) == 4:
        pixel_values = pixel_values.unsqueeze(0)
    b, f, c, h, w = pixel_values.shape
    dev_in = pixel_values.device
    dev = _device_of(pixel_values)
    pers_size = int(h / 2)
    phs = [float(p) for p in ph_list]
    # ---- perspective anchor crops: frame i of clip 0 seen with pitch ph_list[i] ----
    # reference: ((x + 1) / 2 * 255).astype(uint8), evaluated in float32
    x0 = pixel_values[0].to(dev, torch.float32)
    frames = ((x0 + 1) / 2 * 255).permute(0, 2, 3, 1).to(torch.uint8).contiguous()
    maps = [e2p_maps(fov, th, p, pers_size, pers_size, h, w) for p in phs]
    mx = torch.from_numpy(np.stack([m[0] for m in maps])).to(dev)
    my = torch.from_numpy(np.stack([m[1] for m in maps])).to(dev)
    anchor_u8, _ = remap_cubic_wrap(frames, mx, my, paired=True)
    # (anchor_pers / 127.5) - 1 is a float64 numpy expression in the reference
    # (tensor / tensor: torch's tensor / python-scalar path multiplies by the reciprocal, which is not IEEE division)
    div = torch.tensor(127.5, dtype=torch.float64, device=dev)
    anchor_pers = (anchor_u8[:, 0].to(torch.float64) / div - 1).permute(0, 3, 1, 2).unsqueeze(0).expand(b, -1, -1, -1, -1).to(dev_in)
    target = pixel_values.clone()
    # ---- panorama-side visibility masks, largest inscribed rectangle, anchor crop resized to 256x256 ----
    _, _, inside = _cached_p2e(fov, th, phs, pers_size, pers_size, h, w, dev)
    inside_np = inside.cpu().numpy()
    masks, anchors, rel, pitchs = [], [], [], []
    for i in range(f):
        m = torch.from_numpy(1 - inside_np[i].astype(np.int64))[None, None].expand(b, -1, -1, -1).float().to(dev_in)
        masks.append(m)
        ty, tx, rw, rh = get_maxrec_cord(inside_np[i])
        crop = pixel_values[:, i, :, ty:ty + rh, tx:tx + rw]
        anchors.append(F.interpolate(crop, size=(256, 256), mode="bilinear", align_corners=False))
        pitchs.append(torch.tensor([ph_list[i]], device=dev_in))
        rel.append(torch.tensor([int(h / 2 - (ty + ty + rh) / 2), int(w / 2 - (tx + tx + rw) / 2), rh, rw, h, w], device=dev_in))
    pitchs = torch.stack(pitchs, dim=1)
    masks = torch.stack(masks, dim=1)
    anchors = torch.stack(anchors, dim=1)
    rel = torch.stack(rel, dim=0).unsqueeze(0).repeat(b, 1, 1)
    return anchors, anchor_pers, target, masks, rel, pitchs
