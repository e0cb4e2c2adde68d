"""Make the reference's import paths resolve to the native implementation, so ``inference_dual_p2e.py`` runs unchanged.

    import imagine360_b200.dropin as dropin; dropin.install()          # before the script's own imports
    # or:  python -c "import imagine360_b200.dropin as d; d.install(); import runpy; runpy.run_path('inference_dual_p2e.py', run_name='__main__')" --config ...

If the reference tree is importable (``animatediff``, ``src``, ``diffusers`` on sys.path) its modules are imported and
ONLY the symbols on the hot-path boundary are rebound; everything else (CLI helpers, video
I/O, CLIP/SAM glue) keeps coming from the reference.  Without the reference tree, stand-in modules exposing just the
boundary symbols are registered under the same names.
"""
from __future__ import annotations

import importlib
import sys
import types

BOUNDARY = {
    "animatediff.pipelines.pipeline_animation_inference_dual": {"AnimationPipeline": ("imagine360_b200.host.pipeline", "AnimationPipeline")},
    "animatediff.models.unet": {"UNet3DConditionModel": ("imagine360_b200.host.unet3d", "UNet3DConditionModel")},
    "src.models.MVGenModel": {"MultiViewBaseModel": ("imagine360_b200.host.mvgen", "MultiViewBaseModel")},
    "src.utils.Perspective_and_Equirectangular": {"e2p": ("imagine360_b200.dropin", "e2p"), "p2e": ("imagine360_b200.dropin", "p2e")},
    "src.utils.pano": {"pad_pano": ("imagine360_b200.host.geometry", "pad_pano"), "unpad_pano": ("imagine360_b200.host.geometry", "unpad_pano")},
    # geometric pre-processing right before the path: the script's own process_equi /
    # pers2pano_vid reach the GPU through these two classes; get_anchor_target / get_maxrec_cord are rebound whole
    "src.utils.pano_utils.Equirec2Perspec": {"Equirectangular": ("imagine360_b200.host.preprocess", "Equirectangular")},
    "src.utils.pano_utils.Perspec2Equirec": {"Perspective": ("imagine360_b200.host.preprocess", "Perspective")},
    "animatediff.utils.video_mask": {"get_anchor_target": ("imagine360_b200.host.preprocess", "get_anchor_target")},
    "src.modules.utils": {"get_maxrec_cord": ("imagine360_b200.host.preprocess", "get_maxrec_cord")},
    # output side (section 8(f) row 3): uint8 conversion of the decoded video on the GPU
    "animatediff.utils.util": {"save_videos_grid": ("imagine360_b200.host.preprocess", "save_videos_grid")},
    "diffusers": {"AutoencoderKL": ("imagine360_b200.host.vae", "AutoencoderKL"), "DDIMScheduler": ("imagine360_b200.host.ddim", "DDIMScheduler")},
    # conditioning encoders (section 8(f) row 2).  segment_anything is a third-party dependency of the script
    # (inference_dual_p2e.py:369-370): the registry / predictor it imports resolve to the native ViT, with or without the
    # package installed.  CLIPTextModel keeps coming from transformers (the pipeline wraps the instance it is handed).
    "segment_anything": {"sam_model_registry": ("imagine360_b200.host.sam", "sam_model_registry"),
                         "SamPredictor": ("imagine360_b200.host.pipeline", "SamPredictor")},
}


def e2p(e_img, fov_deg, u_deg, v_deg, out_hw, mode=None):
    """Reference signature (src/utils/Perspective_and_Equirectangular/e2p.py:54) over the native gather kernel."""
    from .host import geometry as G
    return G.e2p(e_img, {"FoV": fov_deg, "theta": u_deg, "phi": v_deg}, tuple(out_hw), mode=mode or "bilinear")


def p2e(p_img, fov_deg, u_deg, v_deg, out_hw, mode=None):
    from .host import geometry as G
    return G.p2e(p_img, {"FoV": fov_deg, "theta": u_deg, "phi": v_deg}, tuple(out_hw), mode=mode or "bilinear")


def _ensure_module(name: str):
    try:
        return importlib.import_module(name)
    except Exception:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            sub = ".".join(parts[:i])
            if sub not in sys.modules:
                m = types.ModuleType(sub)
                m.__path__ = []
                sys.modules[sub] = m
                if i > 1:
                    setattr(sys.modules[".".join(parts[: i - 1])], parts[i - 1], m)
        return sys.modules[name]


def install() -> dict:
    """Rebind the boundary symbols; returns {module: [symbols]} for logging."""
    done = {}
    for mod_name, symbols in BOUNDARY.items():
        mod = _ensure_module(mod_name)
        for sym, (src_mod, src_name) in symbols.items():
            setattr(mod, sym, getattr(importlib.import_module(src_mod), src_name))
            done.setdefault(mod_name, []).append(sym)
    return done
