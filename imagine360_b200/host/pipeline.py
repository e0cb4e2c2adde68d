"""Host-side mirror of ``animatediff.pipelines.pipeline_animation_inference_dual.AnimationPipeline``.

Same constructor and ``__call__`` signature / ``video_batch`` contract as the reference (pipeline...dual.py:63-81,
:553-596) so ``inference_dual_p2e.py`` can drive it unchanged.  The denoising loop (:734-809) is exposed on its own as
:meth:`AnimationPipeline.denoise` -- that is the unit ``bench.py`` times: per step one dual-branch
``MultiViewBaseModel.forward`` + one fused CFG/DDIM kernel per branch, no ``empty_cache`` / ``gc`` / ``.item()`` syncs.
"""
from __future__ import annotations

import os
import random
import sys
from dataclasses import dataclass
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from . import geometry as G
from .encoders import SamImageEncoderNative, _DeviceSwitch, wrap_text_encoder
from .unet3d import refresh_packed_weights

BF16 = torch.bfloat16


@dataclass
class AnimationPipelineOutput:
    videos: torch.Tensor


@dataclass
class Conditioning:
    """Step-invariant inputs of the loop (what the reference computes before `for t in timesteps`)."""
    text_pano: torch.Tensor      # [2, 77, D]      uncond (+) cond
    text_pers: torch.Tensor      # [2*m, 77, D]
    feats_pano: torch.Tensor     # [2, F, 4096, C] SAM features (cond == uncond copy, pipeline...dual.py:695)
    feats_pers: torch.Tensor     # [2, m, F, 4096, C]
    rel_pos: torch.Tensor        # [F, 6]
    pitch: torch.Tensor          # [F]
    fps: int = 8


class StepGraph:
    """The dual-branch forward of a denoising step as two CUDA graphs: ``g_adapter`` (SAM features ->
    clean IP tokens of both branches + relative-position tokens; replayed when the clip's conditioning changes) and
    ``g_step`` (one ``MultiViewBaseModel._forward`` on static inputs; replayed every step).  A step is ~1 450 kernel
    launches; replaying them removes the host's launch pace from the loop (at the 16x256x512 single-branch size the eager
    step is host bound).  What varies per step enters through static device buffers: the latents (updated in place by the
    fused CFG + DDIM kernel), the timestep, and the WarpAttn bias of each of the 7 call sites (``BiasSlot``), chosen by the
    same ``random.random() < 0.4`` draws, in the same order, as the eager path; the IP-token noise is drawn inside the
    graph from torch's CUDA generator (graph-safe philox), pano first, like ``MVGenModel.py:186-187``."""

    def __init__(self, pipe, pano_latent, pers_latent, cond, cameras):
        from .mvgen import BiasSlot
        mv = pipe.mv_base_model
        dev = pano_latent.device
        _, m, _, f, ph, pw = pers_latent.shape
        _, _, _, eh, ew = pano_latent.shape
        self.m = m

        def buf(*shape, dtype=BF16):
            return torch.zeros(shape, dtype=dtype, device=dev)

        self.lat_pano, self.lat_pers = buf(1, 4, f, eh, ew), buf(1, m, 4, f, ph, pw)
        self.stat_pano, self.stat_pers = buf(1, 5, f, eh, ew), buf(1, m, 5, f, ph, pw)
        self.text_pano, self.text_pers = buf(*cond.text_pano.shape), buf(*cond.text_pers.shape)
        self.feats_pano = buf(*cond.feats_pano.shape)
        self.feats_pers1 = buf(cond.feats_pers.shape[0], 1, *cond.feats_pers.shape[2:])
        self.rel, self.pitch = buf(2, f, 6, dtype=torch.float32), buf(2, f, dtype=torch.float32)
        self.fps_pano = torch.zeros(2, dtype=torch.int64, device=dev)
        self.fps_pers = torch.zeros(2, m, dtype=torch.int64, device=dev)
        self.t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.cams = dict(zip(("FoV", "theta", "phi"), (list(c[:m]) for c in G.camera_lists(
            {k: (v.reshape(-1) if isinstance(v, torch.Tensor) else v) for k, v in cameras.items() if k in ("FoV", "theta", "phi")}))))
        self.slots = [BiasSlot() for _ in range(7)]
        self._src = {}                      # name -> (source tensor, version) last copied into the static buffers
        self._adapter_dirty = True
        self._load_cond(cond)
        feats_pers = self.feats_pers1.expand(-1, m, -1, -1, -1)
        ntok, dctx = mv.unet.num_tokens, cond.text_pano.shape[-1]
        zero_noise = (buf(2, ntok, dctx), buf(2 * m, ntok, dctx))

        def forward(draws, **kw):
            xin_pano = torch.cat([self.lat_pano, self.stat_pano], dim=1)
            xin_pers = torch.cat([self.lat_pers, self.stat_pers], dim=2)
            return mv._forward(latents=torch.cat([xin_pers] * 2), pano_latent=torch.cat([xin_pano] * 2), timestep=self.t,
                               prompt_embd=self.text_pers, pano_prompt_embd=self.text_pano, cameras=self.cams,
                               use_fps_condition=True, use_ip_plus_cross_attention=True, fps_tensor_pano=self.fps_pano,
                               fps_tensor_pers=self.fps_pers, reference_images_clip_feat_pano=self.feats_pano,
                               reference_images_clip_feat_pers=feats_pers, relative_position_tensor=self.rel,
                               pitchs_tensor=self.pitch, antipodal_draws=draws, **kw)

        # eager warm-up, once per shape: fills every cache the capture relies on (weight packings, both bias variants of all
        # 7 call sites, PE tables, kernel attributes); zero noise is injected so torch's RNG stream is left untouched
        for anti in (False, True):
            for s_ in self.slots:
                s_.antipodal = anti
            tokens = mv._adapter_compute(self.feats_pano, feats_pers, self.rel, self.pitch)
            forward(list(self.slots), ip_noise=zero_noise, adapter_tokens=tokens)
        for s_ in self.slots:
            s_.freeze()
        del tokens
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()            # the eager warm-up's blocks go back before the graphs reserve their own pool
        self.g_adapter, self.g_step = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_adapter):
            self.tokens = mv._adapter_compute(self.feats_pano, feats_pers, self.rel, self.pitch)
        from .. import _lib
        n0 = _lib.LAUNCHES
        with torch.cuda.graph(self.g_step):
            self.pred_pers, self.pred_pano = forward(list(self.slots), adapter_tokens=self.tokens)
        self.kernels_per_replay = _lib.LAUNCHES - n0        # C-ABI launches recorded into the step graph
        # the graphs read the mask / PE / grid tables of this camera set by address: hold them, so that pruning the geometry cache
        # (G.prune_cache) can never free what a live graph reads
        self._geometry = G.cached_entries(self.cams)

    @staticmethod
    def key_of(pipe, pano_latent, pers_latent, cond, cameras):
        cams = G.camera_lists({k: (v.reshape(-1) if isinstance(v, torch.Tensor) else v) for k, v in cameras.items()
                               if k in ("FoV", "theta", "phi")})
        return (id(pipe.mv_base_model), str(pano_latent.device), tuple(pano_latent.shape), tuple(pers_latent.shape),
                tuple(cond.text_pano.shape), tuple(cond.text_pers.shape), tuple(cond.feats_pano.shape), tuple(cond.feats_pers.shape),
                tuple(c[:pers_latent.shape[1]] for c in cams))

    @staticmethod
    def _sig(t):
        return (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.dtype)

    def _copy_if_changed(self, name, dst, src):
        last = self._src.get(name)
        if last is None or last[1] != self._sig(src):
            dst.copy_(src.to(dst.dtype) if src.dtype != dst.dtype else src)
            self._src[name] = (src, self._sig(src))   # holding src pins its storage: (address, version) then means contents
            return True
        return False

    def _load_cond(self, cond):
        self._copy_if_changed("text_pano", self.text_pano, cond.text_pano)
        self._copy_if_changed("text_pers", self.text_pers, cond.text_pers)
        dirty = self._copy_if_changed("feats_pano", self.feats_pano, cond.feats_pano)
        fp = cond.feats_pers
        if fp.shape[1] > 1 and fp.stride(1) != 0:
            raise NotImplementedError("per-view SAM features: the graphed step expects the reference's shared features "
                                      "(pipeline...dual.py:716-717); set I360_CUDA_GRAPH=0")
        dirty |= self._copy_if_changed("feats_pers", self.feats_pers1, fp[:, :1])
        last = self._src.get("rel")
        if last is None or last[1] != self._sig(cond.rel_pos) or self._src["pitch"][1] != self._sig(cond.pitch):
            self.rel.copy_(cond.rel_pos.to(self.rel.device, torch.float32)[None].expand(2, -1, -1))
            self.pitch.copy_(cond.pitch.to(self.rel.device, torch.float32)[None].expand(2, -1))
            self._src["rel"], self._src["pitch"] = (cond.rel_pos, self._sig(cond.rel_pos)), (cond.pitch, self._sig(cond.pitch))
            dirty = True
        if self._src.get("fps") != cond.fps:
            self.fps_pano.fill_(int(cond.fps))
            self.fps_pers.fill_(int(cond.fps))
            self._src["fps"] = cond.fps
        self._adapter_dirty |= dirty

    def load(self, cond, pano_latent, pers_latent, pano_mask, pers_masks, pano_masked, pers_masked):
        self._load_cond(cond)
        if self._adapter_dirty:
            self.g_adapter.replay()
            self._adapter_dirty = False
        self.lat_pano.copy_(pano_latent)
        self.lat_pers.copy_(pers_latent)
        self.stat_pano[:, :1].copy_(pano_mask)
        self.stat_pano[:, 1:].copy_(pano_masked)
        self.stat_pers[:, :, :1].copy_(pers_masks)
        self.stat_pers[:, :, 1:].copy_(pers_masked)

    def step(self, t: int, draws, coeffs, guidance):
        self.t.fill_(int(t))
        for slot, d in zip(self.slots, draws):
            slot.select(bool(d))
        self.g_step.replay()
        sa, sb, sap, sbp = coeffs
        ops.cfg_ddim_step(self.lat_pano, self.pred_pano[0:1], self.pred_pano[1:2], guidance, sa, sb, sap, sbp, out=self.lat_pano)
        ops.cfg_ddim_step(self.lat_pers, self.pred_pers[0:1], self.pred_pers[1:2], guidance, sa, sb, sap, sbp, out=self.lat_pers)


class ResizeLongestSide:
    """``segment_anything.utils.transforms.ResizeLongestSide.apply_image`` restated (third-party package, not vendored
    in the reference tree and absent offline; requirements.txt pins no version): resize a uint8 HxWxC image so its
    longest side is ``target_length`` -- ``int(x * scale + 0.5)`` sizes, PIL bilinear (what torchvision's
    ``resize(to_pil_image(image), size)`` does)."""

    def __init__(self, target_length: int):
        self.target_length = target_length

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int):
        scale = long_side_length * 1.0 / max(oldh, oldw)
        return int(oldh * scale + 0.5), int(oldw * scale + 0.5)

    def apply_image(self, image: np.ndarray) -> np.ndarray:
        from PIL import Image
        h, w = self.get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        return np.array(Image.fromarray(image).resize((w, h), Image.BILINEAR))


class SamPredictor:
    """The three ``segment_anything.SamPredictor`` members the reference pipeline touches (``transform``,
    ``set_torch_image``, ``get_image_embedding``; pipeline...dual.py:171-174,:675-718), restated for boxes without the
    package.  ``sam_model`` is what ``sam_model_registry["vit_b"]()`` returns: ``.image_encoder`` (with ``.img_size``),
    ``.pixel_mean`` / ``.pixel_std`` buffers; ``Sam.preprocess`` (normalise, zero-pad right/bottom to the square
    input) is used when the model has it."""

    def __init__(self, sam_model):
        self.model = sam_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self.features = None
        # a segment_anything ImageEncoderViT is evaluated on the sm_100a kernels when its input is
        # on a GPU (host/encoders.py); any other encoder module is simply called
        enc = sam_model.image_encoder
        self.native_encoder = SamImageEncoderNative(enc) if SamImageEncoderNative.supports(enc) else None

    def _preprocess(self, x):
        if callable(getattr(self.model, "preprocess", None)):
            return self.model.preprocess(x)
        x = (x - self.model.pixel_mean) / self.model.pixel_std
        size = self.model.image_encoder.img_size
        return F.pad(x, (0, size - x.shape[-1], 0, size - x.shape[-2]))

    @torch.no_grad()
    def set_torch_image(self, transformed_image, original_image_size):
        size = self.model.image_encoder.img_size
        assert transformed_image.dim() == 4 and transformed_image.shape[1] == 3 and max(transformed_image.shape[2:]) == size, \
            f"set_torch_image input must be BCHW with long side {size}."
        x = self._preprocess(transformed_image)
        if self.native_encoder is not None and x.is_cuda:
            self.features = self.native_encoder(x)
        else:
            self.features = self.model.image_encoder(x)

    def get_image_embedding(self):
        if self.features is None:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        return self.features


def _make_sam_predictor(image_encoder):
    """The pipeline's predictor (pipeline...dual.py:171-174).  Always this module's mirror of the three members the
    reference touches: it routes a segment_anything ViT through the native encoder, which the package's own
    ``SamPredictor.set_torch_image`` (a plain ``self.model.image_encoder(x)``) would not."""
    return SamPredictor(image_encoder)


class AnimationPipeline:
    def __init__(self, vae, text_encoder, tokenizer, pers_unet, pano_unet, mv_base_model, scheduler, image_encoder=None,
                 image_encoder_name="CLIP"):
        self.vae, self.text_encoder, self.tokenizer = vae, wrap_text_encoder(text_encoder), tokenizer
        self.pers_unet, self.pano_unet, self.mv_base_model, self.scheduler = pers_unet, pano_unet, mv_base_model, scheduler
        self.image_encoder, self.image_encoder_name = image_encoder, image_encoder_name
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8
        self.device = torch.device("cpu")
        # pipeline...dual.py:171-174: the pipeline itself wraps the raw SAM model
        if image_encoder_name == "SAM" and image_encoder is not None and not callable(getattr(image_encoder, "embed_frames", None)):
            self.SAMpredictor = _make_sam_predictor(image_encoder)
            self.SAMProcessor = self.SAMpredictor.transform

    def to(self, device):
        self.device = torch.device(device)
        for m in (self.vae, self.text_encoder, self.mv_base_model, self.image_encoder):
            if isinstance(m, torch.nn.Module) or isinstance(m, _DeviceSwitch):
                m.to(device)
        return self

    @property
    def _execution_device(self):
        return self.device

    def enable_vae_slicing(self):
        if self.vae is not None:
            self.vae.enable_slicing()

    # ------------------------------------------------------------------------------------------------
    def init_noise(self, bs, video_length, equi_h, equi_w, pers_h, pers_w, cameras, device, latents_dtype=BF16, pano_noise=None):
        """pipeline...dual.py:361-387: one panorama noise draw; each view gets it resampled with nearest lookup.
        All frames go through ONE gather launch (the grid only depends on the camera)."""
        cams = {k: v.reshape(-1, *v.shape[2:]) if isinstance(v, torch.Tensor) and v.dim() >= 2 else v for k, v in cameras.items()}
        m = len(G.camera_lists(cams)[0])
        if pano_noise is None:
            pano_noise = torch.randn(bs, video_length, 1, 4, equi_h, equi_w, device=device)
        pano_out = pano_noise.squeeze(2).permute(0, 2, 1, 3, 4).contiguous()                       # b c f h w
        src = pano_noise.squeeze(2).reshape(bs, 1, video_length * 4, equi_h, equi_w).expand(-1, m, -1, -1, -1)
        views = G.e2p(src.reshape(bs * m, video_length * 4, equi_h, equi_w).contiguous(),
                      {k: (list(v) * bs) for k, v in zip(("FoV", "theta", "phi"), G.camera_lists(cams))},
                      (pers_h, pers_w), mode="nearest", grid_dtype=torch.float32)
        pers = views.reshape(bs, m, video_length, 4, pers_h, pers_w).permute(0, 1, 3, 2, 4, 5).contiguous()   # b m c f h w
        return pano_out.to(latents_dtype), pers.to(latents_dtype)

    def prepare_masked_latents_pano(self, video_length, pixels_masked, mask):
        """:427-448  (VAE encode in chunks of 8 frames, posterior sample, x 0.18215; mask nearest-resized)"""
        x = pixels_masked.reshape(-1, *pixels_masked.shape[2:])
        lat = torch.cat([self.vae.encode(x[i:i + 8].to(self.vae.dtype), 8).latent_dist.sample() for i in range(0, x.shape[0], 8)])
        b = pixels_masked.shape[0]
        lat = lat.reshape(b, video_length, *lat.shape[1:]).permute(0, 2, 1, 3, 4) * 0.18215
        mask = mask.transpose(2, 1)
        mask = F.interpolate(mask, size=(mask.shape[2], lat.shape[-2], lat.shape[-1]))
        return lat, mask.to(self.device)

    def prepare_masked_latents_pers(self, video_length, pixels_masked, masks):
        """:451-473"""
        b, f, m = pixels_masked.shape[:3]
        x = pixels_masked.reshape(-1, *pixels_masked.shape[3:])
        lat = torch.cat([self.vae.encode(x[i:i + 8].to(self.vae.dtype), 8).latent_dist.sample() for i in range(0, x.shape[0], 8)])
        lat = lat.reshape(b, f, m, *lat.shape[1:]).permute(0, 2, 3, 1, 4, 5) * 0.18215              # b m c f h w
        masks = masks.permute(0, 3, 1, 2, 4, 5).squeeze(0)
        masks = F.interpolate(masks, size=(m, lat.shape[-2], lat.shape[-1])).unsqueeze(3)
        masks = masks.permute(0, 2, 3, 1, 4, 5)                                                      # b m c f h w
        return lat, masks.to(self.device)

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def denoise(self, pano_latent, pers_latent, pano_mask, pers_masks, pano_masked, pers_masked, cond: Conditioning, cameras,
                num_inference_steps=50, guidance_scale=7.5, on_step=None, step_range=None, inject=None):
        with ops.on_device(pano_latent.device):
            return self._denoise(pano_latent, pers_latent, pano_mask, pers_masks, pano_masked, pers_masked, cond, cameras,
                                 num_inference_steps, guidance_scale, on_step, step_range, inject)

    def _denoise(self, pano_latent, pers_latent, pano_mask, pers_masks, pano_masked, pers_masked, cond: Conditioning, cameras,
                 num_inference_steps=50, guidance_scale=7.5, on_step=None, step_range=None, inject=None):
        """The `for t in timesteps` loop (pipeline...dual.py:734-809) -> (pano_latent, pers_latent).
        ``step_range=(i0, i1)`` runs only loop iterations i0..i1-1 of the ``num_inference_steps`` schedule.
        ``inject(i) -> dict(antipodal_draws=[7 bools], ip_noise=(pano, pers))`` overrides the per-step random draws (tests)."""
        dev = pano_latent.device
        m = pers_latent.shape[1]
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        G.prune_cache()          # bounded when the caller varies the cameras per clip (the reference script never does)
        i0, i1 = step_range if step_range is not None else (0, num_inference_steps)
        if inject is None and dev.type == "cuda" and os.environ.get("I360_CUDA_GRAPH", "1") != "0":
            key = StepGraph.key_of(self, pano_latent, pers_latent, cond, cameras)
            sg = self.__dict__.get("_step_graph")
            if sg is None or sg[0] != key:
                self.__dict__.pop("_step_graph", None)          # one graph (= one activation pool) alive at a time
                del sg
                sg = (key, StepGraph(self, pano_latent, pers_latent, cond, cameras))
                self.__dict__["_step_graph"] = sg
            sg = sg[1]
            sg.load(cond, pano_latent, pers_latent, pano_mask, pers_masks, pano_masked, pers_masked)
            for i, t in list(enumerate(self.scheduler.timesteps_host))[i0:i1]:
                draws = [random.random() < 0.4 for _ in range(7)]           # utils.py:15, same order as the eager path
                sg.step(t, draws, self.scheduler.coefficients(t), guidance_scale)
                if on_step is not None:
                    on_step(i, t)
            return sg.lat_pano.clone(), sg.lat_pers.clone()
        # CFG-doubled per-clip tensors live on the Conditioning object: they are the same objects for every step (and
        # every denoise() call) of a clip, which is what lets the model's adapter cache recognise the clip
        derived = cond.__dict__.get("_derived")
        if derived is None or derived[0] != (dev, m, cond.rel_pos.data_ptr(), cond.rel_pos._version, cond.pitch.data_ptr(),
                                             cond.pitch._version, cond.fps):
            fps_pano = torch.tensor([cond.fps, cond.fps], device=dev)
            derived = ((dev, m, cond.rel_pos.data_ptr(), cond.rel_pos._version, cond.pitch.data_ptr(), cond.pitch._version, cond.fps),
                       fps_pano, fps_pano[:, None].repeat(1, m), cond.rel_pos.to(dev)[None].repeat(2, 1, 1),
                       cond.pitch.to(dev)[None].repeat(2, 1))
            cond.__dict__["_derived"] = derived
        _, fps_pano, fps_pers, rel_pos, pitch = derived
        static_pano = torch.cat([pano_mask.to(BF16), pano_masked.to(BF16)], dim=1)
        static_pers = torch.cat([pers_masks.to(BF16), pers_masked.to(BF16)], dim=2)
        pano_latent, pers_latent = pano_latent.to(BF16).contiguous(), pers_latent.to(BF16).contiguous()
        for i, t in list(enumerate(self.scheduler.timesteps_host))[i0:i1]:
            xin_pano = torch.cat([pano_latent, static_pano], dim=1)
            xin_pers = torch.cat([pers_latent, static_pers], dim=2)
            pred_pers, pred_pano = self.mv_base_model(
                latents=torch.cat([xin_pers] * 2), pano_latent=torch.cat([xin_pano] * 2),
                timestep=torch.tensor([t], device=dev), prompt_embd=cond.text_pers, pano_prompt_embd=cond.text_pano,
                cameras=cameras, use_fps_condition=True, use_ip_plus_cross_attention=True, fps_tensor_pano=fps_pano,
                fps_tensor_pers=fps_pers, reference_images_clip_feat_pano=cond.feats_pano,
                reference_images_clip_feat_pers=cond.feats_pers, relative_position_tensor=rel_pos, pitchs_tensor=pitch,
                **(inject(i) if inject is not None else {}))
            sa, sb, sap, sbp = self.scheduler.coefficients(t)
            pano_latent = ops.cfg_ddim_step(pano_latent, pred_pano[0:1].contiguous(), pred_pano[1:2].contiguous(),
                                            guidance_scale, sa, sb, sap, sbp)
            pers_latent = ops.cfg_ddim_step(pers_latent, pred_pers[0:1].contiguous(), pred_pers[1:2].contiguous(),
                                            guidance_scale, sa, sb, sap, sbp)
            if on_step is not None:
                on_step(i, t)
        return pano_latent, pers_latent

    @torch.no_grad()
    def decode_latents(self, latents, frames_per_call: int = 4):
        """:301-313 -- returns fp32 [b, 3, f, H, W] in [0, 1] on the device (the caller moves it to the host)."""
        b, c, f, h, w = latents.shape
        z = (latents / 0.18215).permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w).to(self.vae.dtype)
        out = torch.cat([self.vae.decode(z[i:i + frames_per_call]).sample for i in range(0, b * f, frames_per_call)])
        out = out.reshape(b, f, *out.shape[1:]).permute(0, 2, 1, 3, 4)
        return (out / 2 + 0.5).clamp(0, 1).float()

    def decode_video(self, pano_latent):
        """pad_pano(latent, 4) -> decode -> crop 32 px (:811-815)"""
        video = self.decode_latents(G.pad_pano(pano_latent, 4))
        return G.unpad_pano(video, 4 * self.vae_scale_factor)

    @torch.no_grad()
    def decode_video_streamed(self, pano_latent, frames_per_call: int = 4, on_frames=None):
        """Output side of the path: ``decode_video`` in chunks of ``frames_per_call`` frames with the
        device -> host copies of chunk k (fp32 video for the reference's return value, uint8 NHWC frames for the mp4
        writer) running on a copy stream UNDER the decode of chunk k+1, and finished uint8 chunks handed to
        ``on_frames(first_frame, frames_u8[n, H, W, 3])`` on the host while the GPU keeps decoding (an encoder thread can
        consume them).  -> (video fp32 [1, 3, f, H, W] in [0, 1], frames uint8 [f, H, W, 3]), both in pinned host memory;
        values are those of ``decode_video(...).cpu()`` and of ``save_videos_grid``'s ``(x * 255).astype(uint8)``
        (animatediff/utils/util.py:55-72) exactly."""
        from .preprocess import frames_to_u8
        dev = pano_latent.device
        with ops.on_device(dev):
            z = G.pad_pano(pano_latent, 4)
            b, c, f, h, w = z.shape
            if b != 1:
                raise NotImplementedError("one clip per call (the reference is batch 1, pipeline...dual.py:598)")
            zf = (z / 0.18215).permute(0, 2, 1, 3, 4).reshape(f, c, h, w).to(self.vae.dtype)
            crop, sf = 4 * self.vae_scale_factor, self.vae_scale_factor
            H, W = h * sf, w * sf - 2 * crop
            video = torch.empty((1, 3, f, H, W), dtype=torch.float32).pin_memory()
            frames = torch.empty((f, H, W, 3), dtype=torch.uint8).pin_memory()
            main, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
            pending = []

            def drain(block_all):
                while pending and (block_all or len(pending) > 1 or pending[0][2].query()):
                    i0, n0, done = pending.pop(0)
                    done.synchronize()
                    if on_frames is not None:
                        on_frames(i0, frames[i0:i0 + n0])

            for i in range(0, f, frames_per_call):
                n = min(frames_per_call, f - i)
                out = self.vae.decode(zf[i:i + n]).sample
                vid = (out / 2 + 0.5).clamp(0, 1).float()[..., crop:crop + W].contiguous()      # [n, 3, H, W]
                u8 = frames_to_u8(vid, back_norm=False)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    frames[i:i + n].copy_(u8, non_blocking=True)
                    video[0, :, i:i + n].copy_(vid.permute(1, 0, 2, 3), non_blocking=True)
                    u8.record_stream(side)
                    vid.record_stream(side)
                    done = torch.cuda.Event()
                    done.record(side)
                pending.append((i, n, done))
                drain(False)
            drain(True)
        return video, frames

    # ------------------------------------------------------------------------------------------------
    def _encode_prompt(self, prompt, device, negative_prompt):
        """:227-299 through the caller-supplied CLIP tokenizer / text encoder -> cat([uncond, cond]) [2 * len(prompt), 77, D]."""
        cfg = getattr(self.text_encoder, "config", None)
        use_mask = bool(getattr(cfg, "use_attention_mask", False))

        def enc(texts):
            inputs = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True,
                                    return_tensors="pt")
            mask = inputs.attention_mask.to(device) if use_mask else None
            return self.text_encoder(inputs.input_ids.to(device), attention_mask=mask)[0]
        negative_prompt = ["" if n is None else n for n in negative_prompt]
        return torch.cat([enc(negative_prompt), enc(prompt)])

    def _sam_features(self, anchor_pixels):
        """pipeline...dual.py:675-718 for one anchor video [f, 3, h, w] in (-1, 1) -> SAM features [f, (h w), c]:
        uint8 conversion, long side resized to the encoder's input size, batches of 8 frames through
        ``set_torch_image`` / ``get_image_embedding``, ``f c h w -> f (h w) c``.  An ``image_encoder`` exposing
        ``embed_frames(frames) -> [f, hw, c]`` (e.g. a pre-computed-feature source) is used directly instead."""
        if callable(getattr(self.image_encoder, "embed_frames", None)):
            return self.image_encoder.embed_frames(anchor_pixels)
        if self.image_encoder_name != "SAM" or not hasattr(self, "SAMpredictor"):
            msg = "imagine360_b200: use_ip_plus_cross_attention needs image_encoder_name='SAM' and a SAM model (or embed_frames)"
            print(msg, file=sys.stderr, flush=True)       # the caller swallows exceptions (inference_dual_p2e.py:596-597)
            raise ValueError(msg)
        image_array = ((anchor_pixels.to(torch.float32) + 1.0) / 2.0 * 255).to(torch.uint8).cpu().numpy().transpose(0, 2, 3, 1)
        frames = torch.stack([torch.as_tensor(self.SAMProcessor.apply_image(np.ascontiguousarray(img)), device=anchor_pixels.device)
                              .permute(2, 0, 1).contiguous() for img in image_array])
        batch = 8
        if frames.shape[0] % batch != 0:
            msg = f"imagine360_b200: the SAM branch needs a multiple of {batch} frames, got {frames.shape[0]} (pipeline...dual.py:685)"
            print(msg, file=sys.stderr, flush=True)
            raise ValueError(msg)
        embeds = []
        for i in range(0, frames.shape[0], batch):
            self.SAMpredictor.set_torch_image(frames[i:i + batch], tuple(frames[0].shape[:2]))
            e = self.SAMpredictor.get_image_embedding()                       # [8, c, h, w]
            embeds.append(e.flatten(2).transpose(1, 2))
        return torch.cat(embeds, dim=0)

    @torch.no_grad()
    def __call__(self, prompt, num_inference_steps=50, guidance_scale_text=7.5, guidance_scale_adapter=7.5, negative_prompt=None,
                 eta=0.0, generator=None, output_type="tensor", return_dict=True, latents_dtype=BF16, video_batch=None,
                 use_outpaint=False, use_ip_plus_cross_attention=False, use_fps_condition=False, ip_plus_condition="image", **kwargs):
        if not (use_outpaint and use_ip_plus_cross_attention and use_fps_condition and ip_plus_condition == "video") \
                or latents_dtype != BF16:
            msg = "imagine360_b200: only the configuration of configs/prompt-dual.yaml (outpaint + video IP adapter + fps, bf16) is on the native path"
            print(msg, file=sys.stderr, flush=True)           # the caller swallows exceptions (inference_dual_p2e.py:596-597)
            raise NotImplementedError(msg)
        device = self._execution_device
        with ops.on_device(device):
            # LoRA merges edit ``weight.data`` in place without bumping the tensor version (inference_dual_p2e.py:193):
            # one content check per clip keeps the packed weights honest
            refresh_packed_weights(self.mv_base_model)
            if isinstance(self.vae, torch.nn.Module):
                refresh_packed_weights(self.vae)
            video = self._generate(prompt, num_inference_steps, guidance_scale_text, negative_prompt, latents_dtype, video_batch, device)
        return AnimationPipelineOutput(videos=video) if return_dict else video

    def _generate(self, prompt, num_inference_steps, guidance_scale_text, negative_prompt, latents_dtype, vb, device):
        pano_px, pano_mask = vb["pano_pixel_values"], vb["pano_mask"]
        pers_px, pers_masks = vb["pers_pixel_values"], vb["pers_masks"]
        cameras, f, m = vb["cameras"], vb["video_length"], pers_px.shape[2]
        pano_px_m = (pano_px.clone() * (pano_mask < 0.5)).to(device)
        pers_px_m = (pers_px.clone() * (pers_masks < 0.5)).to(device)
        # RNG order of the reference: init_noise -> pano VAE samples -> pers VAE samples -> per-step IP noise
        pano_latent, pers_latent = self.init_noise(1, f, vb["pano_H"] // 8, vb["pano_W"] // 8, vb["pers_size"] // 8,
                                                   vb["pers_size"] // 8, cameras, device, latents_dtype)
        pano_masked, pano_mask_l = self.prepare_masked_latents_pano(f, pano_px_m, pano_mask.to(device))
        pers_masked, pers_masks_l = self.prepare_masked_latents_pers(f, pers_px_m, pers_masks.to(device))
        text_pano = self._encode_prompt([prompt], device, [negative_prompt]).to(latents_dtype)
        text_pers = self._encode_prompt([prompt] * m, device, [negative_prompt] * m).to(latents_dtype)
        anchor, anchor_pers = vb["anchor_pixels_values"].to(device), vb["anchor_pixels_values_pers"].to(device)
        assert anchor.shape[0] == 1, "Batch size must be one"                                  # :672
        feats = self._sam_features(anchor[0]).to(latents_dtype)[None]
        feats_p = self._sam_features(anchor_pers[0]).to(latents_dtype)[None]
        cond = Conditioning(text_pano, text_pers, torch.cat([feats, feats]),
                            torch.cat([feats_p, feats_p]).unsqueeze(1).expand(-1, m, -1, -1, -1),
                            vb["relative_position"].to(device).reshape(f, 6), vb["pitchs"].to(device).reshape(f), vb["fps"])
        pano_latent, _ = self.denoise(pano_latent, pers_latent, pano_mask_l, pers_masks_l, pano_masked, pers_masked, cond, cameras,
                                      num_inference_steps, guidance_scale_text)
        video, frames_u8 = self.decode_video_streamed(pano_latent)
        video._i360_frames_u8 = frames_u8      # the drop-in save_videos_grid writes these instead of re-converting fp32
        return video


# ------------------------------------------------------------------------------------------------------
# synthetic inputs of SURVEY.md §8(d) (there are no datasets / checkpoints offline)
# ------------------------------------------------------------------------------------------------------
def synthetic_inputs(frames=16, pano_hw=(512, 1024), views=20, ctx_dim=1024, sam_dim=256, device="cuda", seed=996995,
                     cameras=None):
    g = torch.Generator(device=device).manual_seed(seed)
    H, W = pano_hw
    eh, ew, ph = H // 8, W // 8, H // 16

    def rn(*shape, scale=1.0):
        return (torch.randn(*shape, device=device, generator=g) * scale).to(BF16)

    pano_latent, pers_latent = rn(1, 4, frames, eh, ew), rn(1, views, 4, frames, ph, ph)
    pano_mask = torch.ones(1, 1, frames, eh, ew, device=device, dtype=BF16)
    pano_mask[..., eh // 4: eh // 4 + eh // 2, ew // 2 - eh // 4: ew // 2 + eh // 4] = 0     # centred (H/2)x(H/2) known square
    pers_masks = torch.ones(1, views, 1, frames, ph, ph, device=device, dtype=BF16)
    cond = Conditioning(rn(2, 77, ctx_dim), rn(2 * views, 77, ctx_dim), rn(2, frames, 4096, sam_dim),
                        rn(2, 1, frames, 4096, sam_dim).expand(-1, views, -1, -1, -1),
                        torch.tensor([1.0, 1.0, H / 2 - 1, H / 2 - 1, H, W], device=device)[None].repeat(frames, 1),
                        torch.zeros(frames, device=device), 8)
    if cameras is None:
        cameras = G.get_cameras(90, H // 2, device=device)
    return dict(pano_latent=pano_latent, pers_latent=pers_latent, pano_mask=pano_mask, pers_masks=pers_masks,
                pano_masked=rn(1, 4, frames, eh, ew, scale=0.18215), pers_masked=rn(1, views, 4, frames, ph, ph, scale=0.18215),
                cond=cond, cameras=cameras)


def random_init_(module, seed=0, std=0.02):
    """Default init, then re-randomise the zero-initialised tensors (SURVEY.md trap 4) so every kernel sees data."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.numel() > 0 and p.is_floating_point() and float(p.detach().abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return module
