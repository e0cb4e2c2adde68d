"""Equirectangular <-> perspective geometry of the product path: sampling grids (host, float64, batched over
cameras, cached), resampling through the native gather kernel, WarpAttn's soft masks and spherical PE tables.

Mirrors src/utils/Perspective_and_Equirectangular/{e2p,p2e}.py, src/utils/pano.py:35-99 and src/utils/utils.py of
the reference.  Everything here is step-invariant, so it is built once per (level, cameras, variant) and cached --
the reference rebuilds both mask variants 7x per denoising step.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops

BF16 = torch.bfloat16


def icosahedron_sample_camera():
    """src/utils/pano.py:35-72 -> (theta, phi) in radians for the 20 faces."""
    r_circ = math.sin(2 * math.pi / 5.0)
    r_in = math.sqrt(3) / 12.0 * (3 + math.sqrt(5))
    r_mid = math.cos(math.pi / 5.0)
    step = 2.0 * math.pi / 5.0
    top = math.pi / 2 - math.acos(r_in / r_circ)
    mid = top - 2 * math.acos(r_in / r_mid)
    k = np.arange(5)
    theta = np.concatenate([-np.pi + step / 2 + k * step, -np.pi + step / 2 + k * step, -np.pi + k * step, -np.pi + k * step])
    phi = np.concatenate([np.full(5, top), np.full(5, mid), np.full(5, -mid), np.full(5, -top)])
    return theta, phi


def get_cameras(fov=90, pers_resolution=512, device="cuda"):
    """inference_dual_p2e.py:79-110 (K/R omitted keys are never read on the denoising path)."""
    th, ph = icosahedron_sample_camera()
    th, ph = np.rad2deg(th), np.rad2deg(ph)
    cams = {"height": np.full_like(th, pers_resolution, dtype=int), "width": np.full_like(th, pers_resolution, dtype=int),
            "FoV": np.full_like(th, fov, dtype=int), "theta": th, "phi": ph}
    return {k: torch.from_numpy(v).unsqueeze(0).to(device) for k, v in cams.items()}


def camera_lists(cameras):
    """{'FoV','theta','phi'} tensors/lists of any shape -> flat python float tuples (hashable cache key)."""
    out = []
    for key in ("FoV", "theta", "phi"):
        v = cameras[key]
        if isinstance(v, torch.Tensor):
            v = v.detach().reshape(-1).cpu().double().tolist()
        out.append(tuple(float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)))
    # the reference accepts a scalar for any of the three (index_list_or_scalar,
    # src/utils/Perspective_and_Equirectangular/utils.py:18-23): broadcast it over the cameras
    n = max(len(o) for o in out)
    if any(len(o) not in (1, n) for o in out):
        raise ValueError(f"camera lists of different lengths: {[len(o) for o in out]}")
    return tuple(o * n if len(o) == 1 and n > 1 else o for o in out)


def _rot(axis, angle):
    """Batched Rodrigues: axis [m,3] (not nec. unit), angle [m] -> [m,3,3]."""
    rvec = axis * angle[:, None]
    th = np.linalg.norm(rvec, axis=1)
    safe = np.where(th < 1e-12, 1.0, th)
    k = rvec / safe[:, None]
    K = np.zeros((len(th), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    R = np.eye(3)[None] + np.sin(th)[:, None, None] * K + (1 - np.cos(th))[:, None, None] * (K @ K)
    R[th < 1e-12] = np.eye(3)
    return R


def _rotations(theta, phi):
    m = len(theta)
    z = np.tile(np.array([0.0, 0.0, 1.0]), (m, 1))
    y = np.tile(np.array([0.0, 1.0, 0.0]), (m, 1))
    R1 = _rot(z, np.radians(theta))
    R2 = _rot(np.einsum("mij,mj->mi", R1, y), np.radians(-phi))
    return R1, R2


def pers_lonlat(cams, h, w):
    """lon/lat [m,h,w] (radians) seen by each perspective pixel (map_pers_coords_to_equi, e2p.py:9-36)."""
    fov, theta, phi = (np.asarray(c, dtype=np.float64) for c in cams)
    m = len(theta)
    w_len = np.tan(np.radians(fov / 2.0))
    h_len = np.tan(np.radians(float(h) / w * fov / 2.0))
    ys = np.linspace(-1, 1, w)[None, None, :] * w_len[:, None, None] * np.ones((1, h, 1))
    zs = -(np.linspace(-1, 1, h)[None, :, None] * h_len[:, None, None]) * np.ones((1, 1, w))
    xyz = np.stack([np.ones((m, h, w)), ys, zs], axis=-1)
    xyz /= np.linalg.norm(xyz, axis=-1, keepdims=True)
    R1, R2 = _rotations(theta, phi)
    xyz = np.einsum("mij,mhwj->mhwi", R2 @ R1, xyz)
    return np.arctan2(xyz[..., 1], xyz[..., 0]), -np.arcsin(xyz[..., 2])


def e2p_pixel_grid(cams, eh, ew, ph, pw):
    """Equirect pixel coordinates [m,ph,pw] (x, y) sampled by each perspective pixel (e2p.py:39-51)."""
    lon, lat = pers_lonlat(cams, ph, pw)
    cx, cy = (ew - 1) / 2.0, (eh - 1) / 2.0
    return lon / np.pi * 180 / 180 * cx + cx, lat / np.pi * 180 / 90 * cy + cy


def p2e_pixel_grid(cams, ph, pw, eh, ew, theta_offset=0.0):
    """Perspective pixel coordinates [m,eh,ew] (x, y) sampled by each equirect pixel + validity (p2e.py:9-49)."""
    fov, theta, phi = (np.asarray(c, dtype=np.float64) for c in cams)
    theta = theta + theta_offset
    w_len = np.tan(np.radians(fov / 2.0))[:, None, None]
    h_len = np.tan(np.radians(float(ph) / pw * fov / 2.0))[:, None, None]
    x, y = np.meshgrid(np.linspace(-180, 180, ew), np.linspace(90, -90, eh))
    xyz = np.stack([np.cos(np.radians(x)) * np.cos(np.radians(y)), np.sin(np.radians(x)) * np.cos(np.radians(y)),
                    np.sin(np.radians(y))], axis=-1)
    R1, R2 = _rotations(theta, phi)
    Rinv = np.linalg.inv(R1) @ np.linalg.inv(R2)
    xyz = np.einsum("mij,hwj->mhwi", Rinv, xyz)
    front = xyz[..., 0] > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        yy, zz = xyz[..., 1] / xyz[..., 0], xyz[..., 2] / xyz[..., 0]
    inside = (-w_len < yy) & (yy < w_len) & (-h_len < zz) & (zz < h_len)
    gx = np.where(inside, (yy + w_len) / 2 / w_len * pw, 0)
    gy = np.where(inside, (-zz + h_len) / 2 / h_len * ph, 0)
    return gx, gy, inside & front


def _normalised_grid(px, py, src_h, src_w, dtype, device):
    """kornia.remap's pixel -> [-1,1] normalisation carried out in the image dtype, as the reference does after
    casting the float64 maps to that dtype (e2p.py:74-77); returned as fp32 for the gather kernel."""
    gx = torch.from_numpy(px).to(device).to(dtype)
    gy = torch.from_numpy(py).to(device).to(dtype)
    gx = 2.0 * gx / (src_w - 1) - 1.0
    gy = 2.0 * gy / (src_h - 1) - 1.0
    return torch.stack([gx, gy], dim=-1).float().contiguous()


_grid_cache: dict = {}        # key[1] is always the camera set (camera_lists(...)) the entry was built for
_set_last_use: dict = {}      # camera set -> tick of its last use (prune_cache drops the least recently used sets)
_tick = [0]


def _touch(cams):
    _tick[0] += 1
    _set_last_use[cams] = _tick[0]


def cached_entries(cameras) -> list:
    """Every cached table built for this camera set.  A captured step graph keeps this list: the graph reads the tables by
    address, so they must outlive any pruning of the cache."""
    cams = camera_lists(cameras)
    return [v for k, v in _grid_cache.items() if k[1] == cams]


def prune_cache(max_sets: int = 4) -> None:
    """The reference script uses one fixed icosahedron camera set (inference_dual_p2e.py:79-110) and the tables of a set are
    ~100 MB at 16x512x1024; a caller that varies the cameras per clip would otherwise grow the cache without bound.  Keeps the
    ``max_sets`` most recently used camera sets."""
    sets = {k[1] for k in _grid_cache}
    if len(sets) <= max_sets:
        return
    keep = set(sorted(sets, key=lambda c: _set_last_use.get(c, 0))[-max_sets:])
    for k in [k for k in _grid_cache if k[1] not in keep]:
        del _grid_cache[k]
    for c in [c for c in _set_last_use if c not in keep]:
        del _set_last_use[c]


def e2p(e_img, cameras, out_hw, mode="bilinear", grid_dtype=None):
    """e_img [m, c, eh, ew] -> [m, c, ph, pw] via the native gather kernel (e2p.py:54-77)."""
    cams = camera_lists(cameras)
    m, c, eh, ew = e_img.shape
    gd = grid_dtype or e_img.dtype
    key = ("e2p", cams, eh, ew, out_hw, gd, str(e_img.device))
    _touch(cams)
    if key not in _grid_cache:
        px, py = e2p_pixel_grid(cams, eh, ew, *out_hw)
        _grid_cache[key] = _normalised_grid(px, py, eh, ew, gd, e_img.device)
    out = ops.grid_sample(e_img.float(), _grid_cache[key], nearest=(mode == "nearest"))
    return out.to(e_img.dtype)


def p2e(p_img, cameras, out_hw, mode="bilinear", theta_offset=0.0, grid_dtype=None):
    cams = camera_lists(cameras)
    m, c, ph, pw = p_img.shape
    gd = grid_dtype or p_img.dtype
    key = ("p2e", cams, ph, pw, out_hw, theta_offset, gd, str(p_img.device))
    _touch(cams)
    if key not in _grid_cache:
        px, py, mask = p2e_pixel_grid(cams, ph, pw, *out_hw, theta_offset=theta_offset)
        _grid_cache[key] = (_normalised_grid(px, py, ph, pw, gd, p_img.device),
                            torch.from_numpy(mask[:, None]).to(p_img.device))
    grid, mask = _grid_cache[key]
    out = ops.grid_sample(p_img.float(), grid, nearest=(mode == "nearest")) * mask
    return out.to(p_img.dtype), mask


def pad_pano(x, padding):
    """src/utils/pano.py:75-92 on the last dim."""
    if padding <= 0:
        return x
    return torch.cat([x[..., -padding:], x, x[..., :padding]], dim=-1)


def unpad_pano(x, padding):
    return x if padding <= 0 else x[..., padding:-padding]


# ------------------------------------------------------------------------------------------------------
# WarpAttn masks (src/utils/utils.py:12-142) -> dense additive attention biases
# ------------------------------------------------------------------------------------------------------
def _blur5(x, circular):
    k = torch.arange(5, dtype=x.dtype, device=x.device) - 2
    g = torch.exp(-k.pow(2.0) / 2.0)
    g = g / g.sum()
    if circular:
        x = pad_pano(x, 2)
    y = F.pad(x, (2, 2, 2, 2), mode="replicate")
    y = F.conv2d(F.conv2d(y, g.view(1, 1, 1, 5)), g.view(1, 1, 5, 1))
    return unpad_pano(y, 2) if circular else y


def _one_hot(m, h, w, device, shift_w=0):
    px = torch.zeros((m, h * w, h, w), dtype=torch.float32, device=device)
    idx = torch.arange(h * w, device=device)
    px[:, idx, idx // w, (idx % w + shift_w) % w] = 1.0
    return px


def warp_biases(ph, pw, eh, ew, cameras, device, antipodal: bool, grid_dtype=BF16):
    """-> (bias_equi_q [eh*ew, m*ph*pw], bias_pers_q [m*ph*pw, eh*ew]) bf16: the ``mask[0]`` the reference feeds to
    xformers as attn_bias in the two WarpAttn directions (attn_perspano.py:72-92), values in [-1, 1].

    Sampling grids are quantised to ``grid_dtype`` exactly as the reference's ``.type(image dtype)`` does (the masks are
    built in the activation dtype, bf16 in production: attn_perspano.py:40); interpolation, the "missing pixel" fix,
    the 5x5 Gaussian and the max-normalisation then run in fp32 (the reference's own bf16 arithmetic there depends on
    ATen's bf16 grid_sample/conv kernels and is not pinned by anything)."""
    cams = camera_lists(cameras)
    key = ("bias", cams, ph, pw, eh, ew, antipodal, grid_dtype, str(device))
    _touch(cams)
    if key in _grid_cache:
        return _grid_cache[key]
    m = len(cams[0])
    sh = ew // 2 if antipodal else 0
    probe = torch.empty(0, dtype=grid_dtype, device=device)
    pers_masks = e2p(_one_hot(m, eh, ew, device, sh), cameras, (ph, pw), grid_dtype=probe.dtype).reshape(m, eh, ew, ph, pw)
    equi_masks = p2e(_one_hot(m, ph, pw, device), cameras, (eh, ew), theta_offset=180.0 if antipodal else 0.0,
                     grid_dtype=probe.dtype)[0].reshape(m, ph, pw, eh, ew)
    idx = torch.arange(eh * ew, device=device)
    ei, ej = idx // ew, (idx % ew + sh) % ew
    pers_masks[:, ei, ej] += equi_masks[:, :, :, ei, ej].permute(0, 3, 1, 2)
    pers_masks = pers_masks.clamp(0, 1)
    pidx = torch.arange(ph * pw, device=device)
    pi, pj = pidx // pw, pidx % pw
    equi_masks[:, pi, pj] += pers_masks[:, :, :, pi, pj].permute(0, 3, 1, 2)
    equi_masks = equi_masks.clamp(0, 1)
    pm = _blur5(pers_masks.reshape(m * eh * ew, 1, ph, pw), False)
    em = _blur5(equi_masks.reshape(m * ph * pw, 1, eh, ew), True)

    def norm(x):
        mx = torch.amax(x, dim=(1, 2, 3), keepdim=True)
        mx[mx == 0] = 1.0
        return x / mx * 2 - 1

    pm = norm(pm).reshape(m, eh, ew, ph, pw).permute(1, 2, 0, 3, 4).reshape(eh * ew, m * ph * pw)
    em = norm(em).reshape(m * ph * pw, eh * ew)
    out = (pm.to(BF16).contiguous(), em.to(BF16).contiguous())
    _grid_cache[key] = out
    return out


def spherical_pe_tables(freq_bands, ph, pw, eh, ew, cameras, device, dtype=BF16):
    """SphericalPE of get_coords (utils.py:145-164, transformer.py:190-205) evaluated in the activation dtype, as in
    the reference: -> (pers_pe [m*ph*pw, C], equi_pe [eh*ew, C])."""
    cams = camera_lists(cameras)
    key = ("pe", cams, ph, pw, eh, ew, dtype, str(device), freq_bands.data_ptr(), freq_bands._version)
    _touch(cams)
    if key in _grid_cache:
        return _grid_cache[key]
    x, y = np.meshgrid(np.linspace(-np.pi, np.pi, ew), np.linspace(np.pi / 2, -np.pi / 2, eh))
    equi = torch.tensor(np.stack([x, y], -1), device=device, dtype=dtype)                     # [eh, ew, 2]
    lon, lat = pers_lonlat(cams, ph, pw)
    pers = torch.tensor(np.stack([lon, lat], -1), device=device, dtype=dtype)                  # [m, ph, pw, 2]
    fb = freq_bands.to(device=device, dtype=dtype)

    def enc(c):
        e = c.reshape(-1, 2, 1) * fb
        return torch.cat([torch.sin(e), torch.cos(e)], dim=1).reshape(c.shape[0] if c.dim() == 2 else -1, -1)

    pe_p = enc(pers.reshape(-1, 2)).to(BF16).contiguous()
    pe_e = enc(equi.reshape(-1, 2)).to(BF16).contiguous()
    _grid_cache[key] = (pe_p, pe_e)
    return pe_p, pe_e
