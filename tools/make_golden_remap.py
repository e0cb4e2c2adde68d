"""Golden vectors for the geometric pre-processing row, written from the REAL cv2.remap and the
UNMODIFIED reference classes / functions in this container:  python tools/make_golden_remap.py
-> tests/golden/remap_golden.npz (a few hundred KB).  Nothing here is imported at test time."""
import ast
import importlib.util
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def function_source(rel, fn):
    src = open(os.path.join(REF, rel)).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == fn:
            return ast.get_source_segment(src, node)
    raise KeyError(fn)


E2P = load("E2P", "src/utils/pano_utils/Equirec2Perspec.py")
P2E = load("P2E", "src/utils/pano_utils/Perspec2Equirec.py")
rng = np.random.default_rng(20260117)
out = {"cv2_version": np.array(cv2.__version__), "numpy_version": np.array(np.__version__)}

# 1. raw cv2.remap: random image, maps reaching far outside the image (wrap) and exact half-steps of the 1/32 grid
img = rng.integers(0, 256, (40, 64, 3), dtype=np.uint8)
mx = (rng.random((30, 50)) * 104 - 20).astype(np.float32)
my = (rng.random((30, 50)) * 80 - 20).astype(np.float32)
mx[0, :10] = np.arange(10, dtype=np.float32) / 64 + 3        # ties of round-half-even
my[0, :10] = 5 + np.arange(10, dtype=np.float32) / 64
out.update(raw_img=img, raw_mx=mx, raw_my=my, raw_out=cv2.remap(img, mx, my, cv2.INTER_CUBIC, borderMode=cv2.BORDER_WRAP))

# 2. Equirectangular.GetPerspective / Perspective.GetEquirec of the reference
pano = rng.integers(0, 256, (32, 64, 3), dtype=np.uint8)
views = [(0.0, 0.0), (36.0, 26.565), (-108.0, -52.62), (180.0, 90.0)]
out.update(pano=pano, e2p_views=np.array(views),
           e2p_out=np.stack([E2P.Equirectangular(pano).GetPerspective(90, t, p, 16, 16) for t, p in views]))
pers = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
p2e_phis = [0.0, 12.5, -33.25]
res = [P2E.Perspective(pers, 90, 0, p).GetEquirec(32, 64) for p in p2e_phis]
out.update(pers=pers, p2e_phis=np.array(p2e_phis), p2e_out=np.stack([r[0] for r in res]), p2e_mask=np.stack([r[1] for r in res]))

# 3. process_equi (inference_dual_p2e.py:113) and get_maxrec_cord / get_anchor_target, executed from the reference source
from einops import rearrange  # noqa: E402

ns = {"np": np, "torch": torch, "E2P": E2P, "P2E": P2E, "rearrange": rearrange, "F": torch.nn.functional}
exec(function_source("inference_dual_p2e.py", "process_equi"), ns)
exec(function_source("src/modules/utils.py", "get_maxrec_cord"), ns)
exec(function_source("animatediff/utils/video_mask.py", "get_anchor_target"), ns)
vid = torch.from_numpy(rng.random((2, 3, 32, 64)).astype(np.float32) * 2 - 1)
th = np.array([[0.0, 72.0, -144.0]]); ph = np.array([[0.0, 26.565, -52.62]])
out.update(pe_vid=vid.numpy(), pe_th=th, pe_ph=ph, pe_out=ns["process_equi"](vid, th, ph, pers_resolution=16).numpy())
msk = torch.from_numpy((rng.random((2, 1, 32, 64)) < 0.5).astype(np.uint8))
out.update(pe_mask_in=msk.numpy(), pe_mask_out=ns["process_equi"](msk.repeat(1, 3, 1, 1), th, ph, pers_resolution=16, back_norm=False).numpy())
m = (rng.random((24, 40)) < 0.85).astype(np.int64)
out.update(maxrec_mask=m, maxrec_out=np.array([int(v) for v in ns["get_maxrec_cord"](m)]))
pv = torch.from_numpy(rng.random((3, 3, 64, 128)).astype(np.float32) * 2 - 1)
phl = [0.0, 7.5, -11.25]
a, ap, tg, mk, rel, pit = ns["get_anchor_target"](pv, phl)
out.update(at_in=pv.numpy(), at_ph=np.array(phl), at_anchor_s8=a.numpy()[..., ::8, ::8].copy(),   # every 8th pixel keeps the file small
            at_anchor_pers=ap.numpy(), at_masks=mk.numpy(), at_rel=rel.numpy(),
           at_pitch=pit.numpy())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "remap_golden.npz"), **out)
print({k: getattr(v, "shape", None) for k, v in out.items()})
