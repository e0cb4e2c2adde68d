"""Oracle: equirectangular <-> perspective geometry, WarpAttn masks, spherical PE, circular pad.

Follows src/utils/Perspective_and_Equirectangular/{e2p,p2e}.py, src/utils/pano.py,
src/utils/utils.py and src/modules/transformer.py:170-205 of the reference.  kornia (unpinned,
absent) is restated from its public semantics: remap == grid_sample on pixel coordinates
normalised as 2x/(W-1)-1 (align_corners=True, zeros padding) IN THE IMAGE DTYPE -- the reference
casts the float64 sampling grids to the image dtype before the call (e2p.py:74-75, p2e.py:67-68).
cv2.Rodrigues is replaced by the closed-form axis-angle rotation (same matrix to fp64 rounding).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------
# cameras (inference_dual_p2e.py:79-110, src/utils/pano.py:35-72)
# ------------------------------------------------------------------------------------------
def icosahedron_cameras():
    r_circ = np.sin(2 * np.pi / 5.0)
    r_in = np.sqrt(3) / 12.0 * (3 + np.sqrt(5))
    r_mid = np.cos(np.pi / 5.0)
    step = 2.0 * np.pi / 5.0
    top = np.pi / 2 - np.arccos(r_in / r_circ)
    mid = np.pi / 2.0 - np.arccos(r_in / r_circ) - 2 * np.arccos(r_in / r_mid)
    thetas, phis = [], []
    for i in range(20):
        k = i % 5
        if i < 5:
            th, ph = -np.pi + step / 2.0 + k * step, top
        elif i < 10:
            th, ph = -np.pi + step / 2.0 + k * step, mid
        elif i < 15:
            th, ph = -np.pi + k * step, -mid
        else:
            th, ph = -np.pi + k * step, -top
        thetas.append(th)
        phis.append(ph)
    return np.rad2deg(np.array(thetas)), np.rad2deg(np.array(phis))


def default_cameras(fov=90):
    """theta/phi/FoV of the 20 views, as python floats/ints (what ``.item()`` yields in the reference;
    FoV is an int array there: np.full_like(thetas, fov, dtype=int))."""
    th, ph = icosahedron_cameras()
    return dict(FoV=[int(fov)] * 20, theta=[float(t) for t in th], phi=[float(p) for p in ph])


def _rodrigues(axis, angle):
    """Rotation matrix for rotation vector axis*angle (what cv2.Rodrigues returns)."""
    rvec = np.asarray(axis, np.float64) * angle
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _rotations(theta, phi):
    y_axis = np.array([0.0, 1.0, 0.0], np.float32)
    z_axis = np.array([0.0, 0.0, 1.0], np.float32)
    R1 = _rodrigues(z_axis, np.radians(theta))
    R2 = _rodrigues(np.dot(R1, y_axis), np.radians(-phi))
    return R1, R2


def pers_coords_to_equi(wfov, theta, phi, h, w):
    """map_pers_coords_to_equi (e2p.py:9-36): lon/lat (radians) seen by each perspective pixel."""
    hfov = float(h) / w * wfov
    w_len = np.tan(np.radians(wfov / 2.0))
    h_len = np.tan(np.radians(hfov / 2.0))
    x_map = np.ones([h, w], np.float32)
    y_map = np.tile(np.linspace(-w_len, w_len, w), [h, 1])
    z_map = -np.tile(np.linspace(-h_len, h_len, h), [w, 1]).T
    D = np.sqrt(x_map ** 2 + y_map ** 2 + z_map ** 2)
    xyz = np.stack((x_map, y_map, z_map), axis=2) / D[:, :, None]
    R1, R2 = _rotations(theta, phi)
    xyz = xyz.reshape([h * w, 3]).T
    xyz = np.dot(R2, np.dot(R1, xyz)).T
    lat = np.arcsin(xyz[:, 2])
    lon = np.arctan2(xyz[:, 1], xyz[:, 0])
    return lon.reshape([h, w]), -lat.reshape([h, w])


def pers_pix_to_equi(eh, ew, fov, theta, phi, h, w):
    """map_pers_pix_to_equi (e2p.py:39-51): equirect pixel coordinates sampled by each pers pixel."""
    lon, lat = pers_coords_to_equi(fov, theta, phi, h, w)
    cx, cy = (ew - 1) / 2.0, (eh - 1) / 2.0
    lon = lon / np.pi * 180
    lat = lat / np.pi * 180
    return lon / 180 * cx + cx, lat / 90 * cy + cy


def equi_pix_to_pers(ph, pw, wfov, theta, phi, h, w):
    """map_equi_pix_to_pers (p2e.py:9-49): perspective pixel coordinates sampled by each equirect pixel,
    plus the validity mask."""
    hfov = float(ph) / pw * wfov
    w_len = np.tan(np.radians(wfov / 2.0))
    h_len = np.tan(np.radians(hfov / 2.0))
    x, y = np.meshgrid(np.linspace(-180, 180, w), np.linspace(90, -90, h))
    xyz = np.stack((np.cos(np.radians(x)) * np.cos(np.radians(y)), np.sin(np.radians(x)) * np.cos(np.radians(y)),
                    np.sin(np.radians(y))), axis=2)
    R1, R2 = _rotations(theta, phi)
    R1, R2 = np.linalg.inv(R1), np.linalg.inv(R2)
    xyz = xyz.reshape([h * w, 3]).T
    xyz = np.dot(R1, np.dot(R2, xyz)).T.reshape([h, w, 3])
    front = np.where(xyz[:, :, 0] > 0, 1, 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        xyz = xyz / xyz[:, :, 0:1]
    inside = (-w_len < xyz[:, :, 1]) & (xyz[:, :, 1] < w_len) & (-h_len < xyz[:, :, 2]) & (xyz[:, :, 2] < h_len)
    lon_map = np.where(inside, (xyz[:, :, 1] + w_len) / 2 / w_len * pw, 0)
    lat_map = np.where(inside, (-xyz[:, :, 2] + h_len) / 2 / h_len * ph, 0)
    mask = (np.where(inside, 1, 0) * front) > 0
    return lon_map, lat_map, mask


def _cam(cameras, key, i):
    v = cameras[key]
    v = v[i] if hasattr(v, "__len__") else v
    return v.item() if isinstance(v, torch.Tensor) else v


def remap(img, map_x, map_y, mode):
    """kornia.geometry.transform.remap(align_corners=True) restated; the maps arrive in the dtype the reference cast
    them to (the image dtype -- bf16 in production) and are normalised in that dtype."""
    h, w = img.shape[-2:]
    gx = 2.0 * map_x / (w - 1) - 1.0
    gy = 2.0 * map_y / (h - 1) - 1.0
    grid = torch.stack([gx, gy], dim=-1).to(img.dtype)   # gx/gy were computed in the maps' dtype
    return F.grid_sample(img, grid, mode=mode, padding_mode="zeros", align_corners=True)


def e2p(e_img, cameras, out_hw, mode="bilinear", grid_dtype=None):
    """e2p for tensors (e2p.py:54-77).  e_img: [m, c, he, we], one camera per batch item.
    ``grid_dtype`` (default: the image dtype, as in the reference) lets an fp32 evaluation keep the production
    path's bf16 quantisation of the sampling grid."""
    m, c, he, we = e_img.shape
    lons, lats = [], []
    for i in range(m):
        lon, lat = pers_pix_to_equi(he, we, _cam(cameras, "FoV", i), _cam(cameras, "theta", i), _cam(cameras, "phi", i),
                                    out_hw[0], out_hw[1])
        lons.append(lon)
        lats.append(lat)
    gd = grid_dtype or e_img.dtype
    lons = torch.from_numpy(np.stack(lons)).to(e_img.device).type(gd)
    lats = torch.from_numpy(np.stack(lats)).to(e_img.device).type(gd)
    return remap(e_img, lons, lats, mode)


def p2e(p_img, cameras, out_hw, mode="bilinear", theta_offset=0.0, grid_dtype=None):
    """p2e for tensors (p2e.py:52-71). Returns (equi, mask)."""
    m, c, hp, wp = p_img.shape
    lons, lats, masks = [], [], []
    for i in range(m):
        lon, lat, mask = equi_pix_to_pers(hp, wp, _cam(cameras, "FoV", i), _cam(cameras, "theta", i) + theta_offset,
                                          _cam(cameras, "phi", i), out_hw[0], out_hw[1])
        lons.append(lon)
        lats.append(lat)
        masks.append(mask[None])
    gd = grid_dtype or p_img.dtype
    lons = torch.from_numpy(np.stack(lons)).to(p_img.device).type(gd)
    lats = torch.from_numpy(np.stack(lats)).to(p_img.device).type(gd)
    mask = torch.from_numpy(np.stack(masks)).to(p_img.device)
    return remap(p_img, lons, lats, mode) * mask, mask


# ------------------------------------------------------------------------------------------
# circular padding (src/utils/pano.py:75-99)
# ------------------------------------------------------------------------------------------
def pad_pano(x, padding):
    if padding <= 0:
        return x
    return torch.cat([x[..., -padding:], x, x[..., :padding]], dim=-1)


def unpad_pano(x, padding):
    return x if padding <= 0 else x[..., padding:-padding]


# ------------------------------------------------------------------------------------------
# WarpAttn masks (src/utils/utils.py:12-142)
# ------------------------------------------------------------------------------------------
def _gaussian_blur5(x, circular):
    """kornia gaussian_blur2d((5,5),(1,1),'replicate'); the equirect masks are first circular-padded by 2
    and cropped after (utils.py:26-29)."""
    k = torch.arange(5, dtype=x.dtype, device=x.device) - 2
    g = torch.exp(-k.pow(2.0) / 2.0)
    g = g / g.sum()
    if circular:
        x = pad_pano(x, 2)
    y = F.pad(x, (2, 2, 2, 2), mode="replicate")
    y = F.conv2d(y, g.view(1, 1, 1, 5))
    y = F.conv2d(y, g.view(1, 1, 5, 1))
    if circular:
        y = unpad_pano(y, 2)
    return y


def _one_hot_images(m, h, w, dtype, device, shift_w=0):
    """pixels[:, i, j] is an image with a single 1 at (i, (j + shift_w) % w)  (utils.py:52-60, :100-108)."""
    px = torch.zeros((m, h * w, h, w), dtype=dtype, device=device)
    idx = torch.arange(h * w, device=device)
    ii, jj = idx // w, idx % w
    px[:, idx, ii, (jj + shift_w) % w] = 1.0
    return px


def raw_masks(ph, pw, eh, ew, cameras, device, dtype, antipodal: bool, grid_dtype=None):
    """get_masks (utils.py:43-89) / get_oppo_masks (:91-142) before blur/normalisation.
    pers_masks [m, eh, ew, ph, pw], equi_masks [m, ph, pw, eh, ew]."""
    m = len(cameras["FoV"])
    pers_px = _one_hot_images(m, ph, pw, dtype, device)
    equi_px = _one_hot_images(m, eh, ew, dtype, device, shift_w=ew // 2 if antipodal else 0)
    pers_masks = e2p(equi_px, cameras, (ph, pw), grid_dtype=grid_dtype)
    equi_masks = p2e(pers_px, cameras, (eh, ew), theta_offset=180.0 if antipodal else 0.0, grid_dtype=grid_dtype)[0]
    pers_masks = pers_masks.reshape(m, eh, ew, ph, pw)
    equi_masks = equi_masks.reshape(m, ph, pw, eh, ew)
    # "fix missing pixels": add the transposed other-direction correspondences, clamp to [0,1]
    sh = ew // 2 if antipodal else 0
    idx = torch.arange(eh * ew, device=device)
    ei, ej = idx // ew, (idx % ew + sh) % ew
    pm_px = equi_masks[:, :, :, ei, ej].permute(0, 3, 1, 2)        # 'm h w l -> m l h w'
    pers_masks[:, ei, ej] += pm_px
    pers_masks = pers_masks.clamp(0, 1)
    pidx = torch.arange(ph * pw, device=device)
    pi, pj = pidx // pw, pidx % pw
    em_px = pers_masks[:, :, :, pi, pj].permute(0, 3, 1, 2)
    equi_masks[:, pi, pj] += em_px
    equi_masks = equi_masks.clamp(0, 1)
    return pers_masks, equi_masks


def merged_masks(ph, pw, eh, ew, cameras, device, dtype, antipodal: bool, grid_dtype=None):
    """get_merged_masks (utils.py:12-41) for one variant.  The reference computes BOTH variants and picks
    the antipodal one when random.random() < 0.4 (:15-21); the caller owns that draw."""
    pers_masks, equi_masks = raw_masks(ph, pw, eh, ew, cameras, device, dtype, antipodal, grid_dtype)
    m = pers_masks.shape[0]
    pm = _gaussian_blur5(pers_masks.reshape(m * eh * ew, 1, ph, pw), circular=False)
    em = _gaussian_blur5(equi_masks.reshape(m * ph * pw, 1, eh, ew), circular=True)

    def norm(x):
        mx = torch.amax(x, dim=(1, 2, 3), keepdim=True)
        mx[mx == 0] = 1.0
        return x / mx * 2 - 1

    return norm(pm).reshape(m, eh, ew, ph, pw), norm(em).reshape(m, ph, pw, eh, ew)


def polar_coords(ph, pw, eh, ew, cameras, device, dtype):
    """get_coords (utils.py:145-164)."""
    x, y = np.meshgrid(np.linspace(-np.pi, np.pi, ew), np.linspace(np.pi / 2, -np.pi / 2, eh))
    equi = torch.tensor(np.stack([x, y]), device=device, dtype=dtype).permute(1, 2, 0)
    pers = []
    for i in range(len(cameras["FoV"])):
        lon, lat = pers_coords_to_equi(_cam(cameras, "FoV", i), _cam(cameras, "theta", i), _cam(cameras, "phi", i), ph, pw)
        pers.append(torch.tensor(np.stack([lon, lat]), device=device, dtype=dtype))
    return torch.stack(pers, 0).permute(0, 2, 3, 1), equi


def spherical_freqs(n_freqs, device=None):
    """SphericalPE.__init__ (src/modules/transformer.py:170-188)."""
    base = 2 if n_freqs <= 80 else 5000 ** (1 / (n_freqs / 2.5))
    return base ** torch.linspace(0, n_freqs - 1, n_freqs, device=device)


def spherical_pe(coords, freq_bands):
    """SphericalPE.forward (transformer.py:190-205): [..., 2] -> [..., 4*N_freqs] laid out as
    [sin(theta f), sin(phi f), cos(theta f), cos(phi f)] blocks."""
    shape = coords.shape[:-1]
    c = coords.reshape(-1, 2, 1)
    enc = c * freq_bands.to(c.dtype)
    pe = torch.cat([torch.sin(enc), torch.cos(enc)], dim=1)
    return pe.reshape(*shape, -1)
