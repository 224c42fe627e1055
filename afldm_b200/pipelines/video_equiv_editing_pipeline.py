"""Video editing with cross-frame attention on the alias-free SD-1.5 UNet: the call surface and loop structure of
/root/reference/afldm/pipelines/video_equiv_editing_pipeline.py (``image2latent`` :214-226, ``ddim_inversion`` :174-211,
``get_timesteps`` :319-327, ``__call__`` :330-748; driven by scripts/video_editing.py).

Flow of ``__call__`` (:503-748): install ``CrossFrameAttnProcessor`` on every attention; per frame, VAE-encode and
DDIM-invert it with the inversion prompt and NO guidance (frame 0 in STORE state, the others in LOAD state, :593-607);
run the reverse pass of frame 0 with classifier-free guidance in STORE state to record its self-attention inputs
(``save_activations`` :612-649); denoise every frame in LOAD state against those maps (:659-697); decode each frame
(:721-727); restore the processors (:743).

Differences, all on the host side: (1) the CLIP text encoder is outside this build, so the text conditions arrive as
``prompt_embeds`` / ``negative_prompt_embeds`` / ``inv_prompt_embeds`` tensors [1, 77, 768] (or through a user-supplied
``encode_prompt`` callable); (2) the reference evaluates the frames one by one (:668-692, B = 2 per call) - they are
independent given the frame-0 maps, so ``frame_batch`` frames go through the UNet together ([uncond x F | cond x F]; the
attention kernel reads K/V batch b // F, the same repeat the reference materialises, cross_frame_attn.py:91-97);
(3) SDEdit initialisation (``use_sdedit``) is not provided.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import ops, parallel
from ..models.af_vae import AutoencoderKL
from ..models.unet_2d_condition import UNet2DConditionModel
from ..schedulers.ddim import DDIMScheduler
from .cross_frame_attn import AttnState, CrossFrameAttnProcessor

SD15_DDIM = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 steps_offset=1, set_alpha_to_one=False, clip_sample=False, prediction_type="epsilon",
                 timestep_spacing="leading")
SD15_VAE = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=[128, 256, 512, 512],
                layers_per_block=2, scaling_factor=0.18215, sample_size=512)


class StableDiffusionPipelineOutput:
    def __init__(self, images, nsfw_content_detected=None):
        self.images, self.nsfw_content_detected = images, nsfw_content_detected


def _to_tensor(img) -> torch.Tensor:
    """VaeImageProcessor.preprocess for one frame: PIL / uint8 array -> [1,3,H,W] in [-1, 1]; tensors pass ([0,1] -> [-1,1])."""
    if isinstance(img, torch.Tensor):
        t = img if img.ndim == 4 else img.unsqueeze(0)
        return 2.0 * t - 1.0 if float(t.min()) >= 0.0 else t
    arr = np.asarray(img)
    t = torch.from_numpy(arr.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    return 2.0 * t - 1.0


class _CapturedStep:
    """One step  x <- cx x + ce eps(x, t)  of the video loops as a CUDA graph, eps with or without classifier-free
    guidance; t, (cx, ce) and the cross-frame slot index live in device memory (cf. ``ldm_pipeline.GraphedDenoiser``).
    The cross-frame processors must be in slot mode (``AttnState.enable_slots``) and in the state (STORE / LOAD) the
    captured step is going to be replayed in."""

    def __init__(self, unet, frames: int, size: int, cond: torch.Tensor, uncond: Optional[torch.Tensor], guidance: float,
                 warmup: int = 2):
        dev = unet.device
        c = unet.config.in_channels
        self.unet, self.frames, self.guidance = unet, frames, float(guidance)
        self.x = torch.zeros((frames, size, size, c), dtype=torch.float32, device=dev)            # NHWC, updated in place
        nb = frames * (2 if uncond is not None else 1)
        self.t = torch.ones((nb,), dtype=torch.float32, device=dev)
        self.coef = torch.tensor([1.0, 0.0], dtype=torch.float32, device=dev)
        self.x2 = torch.zeros((nb, size, size, c), dtype=torch.float32, device=dev) if uncond is not None else None
        parts = ([uncond.expand(frames, -1, -1)] if uncond is not None else []) + [cond.expand(frames, -1, -1)]
        self.ehs = torch.cat(parts, dim=0).to(device=dev, dtype=torch.float32).contiguous()   # resident before capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.x.zero_()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops._lib.launch_count()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self._body()
        self.launches_per_step = ops._lib.launch_count() - n0

    def set_conditions(self, cond: torch.Tensor, uncond: Optional[torch.Tensor]) -> None:
        """Refresh the text conditions of a step captured earlier (same shapes)."""
        f = self.frames
        if self.x2 is not None:
            self.ehs[:f].copy_(uncond.expand(f, -1, -1))
            self.ehs[f:].copy_(cond.expand(f, -1, -1))
        else:
            self.ehs.copy_(cond.expand(f, -1, -1))

    def _body(self):
        f = self.frames
        if self.x2 is None:
            eps = ops.nhwc(self.unet(ops.nchw_view(self.x), self.t, encoder_hidden_states=self.ehs, return_dict=False)[0])
        else:
            ops.axpby(self.x, self.x, 1.0, 0.0, out=self.x2[:f])          # [uncond x F | cond x F] input
            ops.axpby(self.x, self.x, 1.0, 0.0, out=self.x2[f:])
            e = ops.nhwc(self.unet(ops.nchw_view(self.x2), self.t, encoder_hidden_states=self.ehs, return_dict=False)[0])
            eps = ops.axpby(e[:f], e[f:], 1.0 - self.guidance, self.guidance)
        ops.axpby(self.x, eps, self.coef, None, out=self.x)


class VideoEquivariantEditingPipeline:
    def __init__(self, vae: AutoencoderKL, unet: UNet2DConditionModel, scheduler: DDIMScheduler,
                 encode_prompt: Optional[Callable[[str], torch.Tensor]] = None):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.encode_prompt_fn = encode_prompt
        self.attn_state = AttnState()
        self.replayed_launches = 0                # kernels launched through graph replays (the eager counter misses them)
        self._captured_sig = None
        self._captured = {}                       # captured steps of the CUDA-graph path, per (steps, size, frames, ...)
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8

    @classmethod
    def from_config(cls, unet_config=None, vae_config=None, scheduler_config=None, seed: int = 0, alias_free: bool = True):
        """Randomly initialised pipeline of the SD-1.5 architecture (no checkpoint is reachable offline)."""
        from ..af_modules.af_api import make_af_unet, make_af_vae_from_config
        from ..models.unet_2d_condition import SD15_UNET
        torch.manual_seed(seed)
        unet = UNet2DConditionModel.from_config(unet_config or SD15_UNET)
        vae = AutoencoderKL.from_config(vae_config or SD15_VAE)
        if alias_free:
            make_af_unet(unet)
            make_af_vae_from_config(vae)
        return cls(vae, unet, DDIMScheduler.from_config(scheduler_config or SD15_DDIM))

    def to(self, device):
        self.unet.to(device)
        if self.vae is not None:
            self.vae.to(device)
        return self

    @property
    def device(self):
        return self.unet.device

    # ------------------------------------------------------------------ pieces of the reference pipeline
    @torch.no_grad()
    def image2latent(self, image) -> torch.Tensor:
        """:214-226: posterior MEAN of the VAE encoder, scaled."""
        x = _to_tensor(image).to(device=self.device, dtype=torch.float32)
        return self.vae.encode(x).latent_dist.mean * self.vae.config.scaling_factor

    def get_timesteps(self, num_inference_steps: int, strength: float):
        """:319-327."""
        init_timestep = min(int(num_inference_steps * strength), num_inference_steps)
        t_start = max(num_inference_steps - init_timestep, 0)
        return self.scheduler.timesteps[t_start:], num_inference_steps - t_start

    def _eps(self, latents: torch.Tensor, t: int, cond: torch.Tensor, guidance_scale: float, uncond: Optional[torch.Tensor]):
        """UNet evaluation with classifier-free guidance: batch [uncond x F | cond x F] (:668-691)."""
        f = latents.shape[0]
        if uncond is None:
            return self.unet(latents, t, encoder_hidden_states=cond.expand(f, -1, -1).contiguous()).sample
        x2 = torch.cat([latents, latents], dim=0)
        ehs = torch.cat([uncond.expand(f, -1, -1), cond.expand(f, -1, -1)], dim=0).contiguous()
        e = ops.nhwc(self.unet(x2, t, encoder_hidden_states=ehs).sample)
        eu, ec = e[:f], e[f:]
        # eps = eu + s (ec - eu) = (1 - s) eu + s ec
        return ops.nchw_view(ops.axpby(eu.contiguous(), ec.contiguous(), 1.0 - guidance_scale, guidance_scale))

    @torch.no_grad()
    def ddim_inversion(self, latent: torch.Tensor, timesteps, cond: torch.Tensor, scale: float = 1.0, bar: bool = False,
                       attn_invert: bool = False, uncond: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:174-211: deterministic DDIM run backwards over ``reversed(timesteps)``."""
        sch = self.scheduler
        ts = [int(t) for t in reversed(list(timesteps))]
        latent = latent.to(device=self.device, dtype=torch.float32)
        for i, t in enumerate(ts):
            a_t = sch.alphas_cumprod[t]
            a_prev = sch.alphas_cumprod[ts[i - 1]] if i > 0 else sch.final_alpha_cumprod
            mu, mu_prev = a_t ** 0.5, a_prev ** 0.5
            sigma, sigma_prev = (1 - a_t) ** 0.5, (1 - a_prev) ** 0.5
            if attn_invert:
                self.attn_state.set_timestep(t)
            eps = self._eps(latent, t, cond, scale, uncond if scale != 1.0 else None)
            # latent' = mu * (latent - sigma_prev * eps) / mu_prev + sigma * eps
            latent = ops.nchw_view(ops.axpby(ops.nhwc(latent), ops.nhwc(eps), float(mu / mu_prev),
                                             float(sigma - mu * sigma_prev / mu_prev)))
        return latent

    def _embed(self, text, embeds, what: str) -> torch.Tensor:
        if embeds is None:
            if self.encode_prompt_fn is None or text is None:
                raise ValueError(f"{what}: the CLIP text encoder is outside this build - pass the embedding tensor "
                                 f"[1, 77, {self.unet.config.cross_attention_dim}] or construct the pipeline with encode_prompt=")
            embeds = self.encode_prompt_fn(text)
        return embeds.to(device=self.device, dtype=torch.float32).reshape(1, -1, self.unet.config.cross_attention_dim)

    @torch.no_grad()
    def edit_latents(self, frame_latents: torch.Tensor, pos: torch.Tensor, neg: Optional[torch.Tensor], inv: torch.Tensor,
                     num_inference_steps: int = 50, strength: float = -1, guidance_scale: float = 7.5,
                     frame_batch: Optional[int] = None, start_latents: Optional[torch.Tensor] = None,
                     reference_latent: Optional[torch.Tensor] = None, use_cuda_graph: bool = False,
                     reference_start_latent: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The latent-space part of ``__call__`` (:503-697) on VAE-encoded frames [F,4,h,w]: cross-frame processors,
        per-frame inversion, STORE pass of the reference frame, LOAD denoising of every frame.

        ``reference_latent`` [1,4,h,w]: the frame whose attention maps everybody attends to (default: frame 0 of
        ``frame_latents``).  A rank that holds only a SLICE of the video passes the encoded frame 0 here and re-runs its
        inversion + STORE pass locally - ~1/F of the work - instead of receiving the maps of 2 x 16 attention layers x
        steps from rank 0; with caller-supplied ``start_latents`` it also passes the reference frame's
        start latent as ``reference_start_latent`` so that its STORE pass starts where rank 0's does.
        ``use_cuda_graph``: replay captured steps (``_edit_latents_captured``)."""
        if use_cuda_graph and (frame_batch is None or frame_batch >= frame_latents.shape[0]):
            return self._edit_latents_captured(frame_latents, pos, neg, inv, num_inference_steps, strength, guidance_scale,
                                               start_latents, reference_latent, reference_start_latent)
        dev = self.device
        frame_latents = frame_latents.to(device=dev, dtype=torch.float32)
        num_frames = frame_latents.shape[0]
        own_reference = reference_latent is None
        ref0 = frame_latents[:1] if own_reference else reference_latent.to(device=dev, dtype=torch.float32)
        # cross-frame processors on every attention (:503-512); cross-attention calls pass through them (:126-128)
        ori = self.unet.attn_processors
        st = self.attn_state = AttnState()
        self.unet.set_attn_processor({k: CrossFrameAttnProcessor(st) for k in ori})
        try:
            self.scheduler.set_timesteps(num_inference_steps)
            timesteps = self.scheduler.timesteps
            if strength >= 0:
                timesteps, num_inference_steps = self.get_timesteps(num_inference_steps, strength)
            ts = [int(t) for t in timesteps]

            # 5. per-frame DDIM inversion, the reference frame stores its attention inputs, the others load them (:591-607)
            st.reset()
            lat0 = self.ddim_inversion(ref0, ts, inv, 1.0, attn_invert=True)
            st.to_load()
            fb = frame_batch or num_frames
            rest = frame_latents[1:] if own_reference else frame_latents
            inverted = [lat0] if own_reference else []
            for lo in range(0, rest.shape[0], fb):
                inverted.append(self.ddim_inversion(rest[lo:lo + fb].contiguous(), ts, inv, 1.0, attn_invert=True))
            lat = torch.cat(inverted, dim=0).contiguous()
            if start_latents is not None:           # the reference lets caller-supplied start latents through (:579-589)
                lat = start_latents.to(device=dev, dtype=torch.float32)

            # 6. reverse pass of the reference frame in STORE state: records the maps the other frames attend to (:612-649)
            st.reset()
            st.set_store_id(0)
            x = lat[:1] if own_reference else lat0
            if not own_reference and reference_start_latent is not None:
                x = reference_start_latent.to(device=dev, dtype=torch.float32)
            for t in ts:
                st.set_timestep(t)
                eps = self._eps(x, t, pos, guidance_scale, neg)
                x = self.scheduler.step(eps, t, x).prev_sample

            # 7. denoising loop: every frame in LOAD state (:657-697)
            st.to_load()
            for t in ts:
                st.set_timestep(t)
                eps_parts = [self._eps(lat[lo:lo + fb], t, pos, guidance_scale, neg) for lo in range(0, num_frames, fb)]
                eps = eps_parts[0] if len(eps_parts) == 1 else torch.cat([e.contiguous() for e in eps_parts], dim=0)
                lat = self.scheduler.step(eps, t, lat).prev_sample
        finally:
            self.unet.set_attn_processor(ori)       # :743
        return ops.to_nchw_contiguous(ops.nhwc(lat))

    @torch.no_grad()
    def _edit_latents_captured(self, frame_latents, pos, neg, inv, num_inference_steps, strength, guidance_scale,
                               start_latents, reference_latent, reference_start_latent=None) -> torch.Tensor:
        """``edit_latents`` with every UNet evaluation + update replayed from CUDA graphs: the cross-frame maps live in
        device tables indexed by the step (``AttnState.enable_slots``), timestep / coefficients / slot are device
        scalars refreshed between replays.  Four captured steps: inversion (reference frame: STORE; other frames: LOAD,
        no guidance) and denoising (reference frame: STORE; all frames: LOAD, guidance)."""
        dev = self.device
        frame_latents = frame_latents.to(device=dev, dtype=torch.float32)
        num_frames, size = frame_latents.shape[0], frame_latents.shape[-1]
        own_reference = reference_latent is None
        ref0 = frame_latents[:1] if own_reference else reference_latent.to(device=dev, dtype=torch.float32)
        rest = frame_latents[1:] if own_reference else frame_latents
        sch = self.scheduler
        sch.set_timesteps(num_inference_steps)
        timesteps = sch.timesteps
        if strength >= 0:
            timesteps, _ = self.get_timesteps(num_inference_steps, strength)
        ts = [int(t) for t in timesteps]
        n = len(ts)
        # coefficient tables: inversion runs over reversed(ts) and writes slot n-1-i (the slot of its timestep)
        rev = list(reversed(ts))
        inv_coef = []
        for i, t in enumerate(rev):
            a_t = sch.alphas_cumprod[t]
            a_prev = sch.alphas_cumprod[rev[i - 1]] if i > 0 else sch.final_alpha_cumprod
            mu, mu_prev = a_t ** 0.5, a_prev ** 0.5
            sigma, sigma_prev = (1 - a_t) ** 0.5, (1 - a_prev) ** 0.5
            inv_coef.append([float(mu / mu_prev), float(sigma - mu * sigma_prev / mu_prev)])
        inv_coef = torch.tensor(inv_coef, dtype=torch.float32, device=dev)
        den_coef = torch.stack([torch.stack(sch.coefficients(t)) for t in ts]).to(device=dev, dtype=torch.float32)
        tdev = torch.tensor(ts, dtype=torch.float32, device=dev)

        ori = self.unet.attn_processors
        do_cfg = neg is not None
        key = (n, size, num_frames, rest.shape[0], do_cfg, float(guidance_scale), str(dev))
        # what the captured steps bake in besides the shapes: the storage / version of every parameter (the packed weight
        # copies are rebuilt from them and the graphs hold raw pointers to those copies), the numeric class, the switches
        sig = (tuple((p.data_ptr(), p._version) for p in self.unet.parameters()), ops.default_conv_algo(), ops.F16_CONV,
               ops.F16_ATTENTION, ops.FUSE_GN_PROLOGUE, ops.FUSE_CONCAT, ops.SHORTCUT_SIDE_STREAM)
        if sig != self._captured_sig:
            torch.cuda.synchronize(dev)
            self._captured.clear()
            self._captured_sig = sig
        cached = self._captured.get(key)
        if cached is None:
            st = AttnState()
            st.enable_slots(n, dev)
            # the processors own the device tables the captured graphs address: they live as long as the graphs do
            cached = self._captured[key] = {"state": st, "steps": {},
                                            "procs": {k: CrossFrameAttnProcessor(st) for k in ori}}
            while len(self._captured) > 2:                              # graphs pin their activations: keep two shapes
                torch.cuda.synchronize(dev)                             # nothing of the evicted graphs is still in flight
                self._captured.pop(next(iter(self._captured)))
        st = self.attn_state = cached["state"]
        slots = st.slots
        self.unet.set_attn_processor(cached["procs"])

        def step_for(name, frames, cond, uncond):
            """The captured step ``name``, built (in the cross-frame state it is first needed in) or refreshed."""
            stp = cached["steps"].get(name)
            if stp is None:
                stp = cached["steps"][name] = _CapturedStep(self.unet, frames, size, cond, uncond, guidance_scale)
            else:
                stp.set_conditions(cond.to(dev), None if uncond is None else uncond.to(dev))
            return stp

        def run(step: _CapturedStep, x0, order, coef):
            step.x.copy_(ops.nhwc(x0))
            for i, slot in enumerate(order):
                step.t.copy_(tdev[slot:slot + 1].expand(step.t.shape[0]))      # device-side broadcast of the timestep
                step.coef.copy_(coef[i])
                slots.set_slot(slot)
                step.graph.replay()
            self.replayed_launches += step.launches_per_step * len(order)
            return ops.to_nchw_contiguous(step.x)

        try:
            inv_order = [n - 1 - i for i in range(n)]
            st.reset()                                                  # STORE
            lat0 = run(step_for("inv_store", 1, inv, None), ref0, inv_order, inv_coef)
            st.to_load()
            parts = [lat0] if own_reference else []
            if rest.shape[0] > 0:
                parts.append(run(step_for("inv_load", rest.shape[0], inv, None), rest, inv_order, inv_coef))
            lat = torch.cat(parts, dim=0).contiguous()
            if start_latents is not None:
                lat = start_latents.to(device=dev, dtype=torch.float32)
            st.reset()
            st.set_store_id(0)
            ref_start = lat[:1] if own_reference else lat0
            if not own_reference and reference_start_latent is not None:
                ref_start = reference_start_latent.to(device=dev, dtype=torch.float32)
            run(step_for("den_store", 1, pos, neg if do_cfg else None), ref_start, list(range(n)), den_coef)
            st.to_load()
            lat = run(step_for("den_load", num_frames, pos, neg if do_cfg else None), lat, list(range(n)), den_coef)
        finally:
            self.unet.set_attn_processor(ori)
        return lat

    # ------------------------------------------------------------------ the call
    @torch.no_grad()
    def __call__(self, images: Sequence, prompt: Optional[str] = None, num_inference_steps: int = 50,
                 guidance_scale: float = 7.5, strength: float = -1, negative_prompt: Optional[str] = None,
                 generator=None, latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None, inv_prompt: str = "",
                 inv_prompt_embeds: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, use_sdedit: bool = False, frame_batch: Optional[int] = None,
                 use_cuda_graph: bool = False, shard: Optional[Tuple[int, int]] = None, **kwargs):
        """:330-748.  ``shard = (rank, world_size)``: this process edits frames ``shard_bounds(len(images), rank, world)``
        of the video (it encodes frame 0 as well and re-runs its inversion + STORE pass - ``edit_latents``), decodes
        them, and the call ends with the ONE collective of the path, ``parallel.gather_frames`` of the decoded frames
        (SURVEY.md 8(e)): every rank returns the whole edited video.  Needs an initialised process group."""
        if use_sdedit:
            raise NotImplementedError("use_sdedit: the inversion-based initialisation is the path the reference's script uses")
        dev = self.device
        num_frames = len(images)
        do_cfg = guidance_scale > 1.0
        pos = self._embed(prompt, prompt_embeds, "prompt")
        neg = self._embed(negative_prompt if negative_prompt is not None else "", negative_prompt_embeds, "negative prompt") \
            if do_cfg else None
        inv = self._embed(inv_prompt, inv_prompt_embeds, "inversion prompt")

        lo, hi = (0, num_frames) if shard is None else parallel.shard_bounds(num_frames, int(shard[0]), int(shard[1]))
        frame_latents = torch.cat([self.image2latent(img) for img in images[lo:hi]], dim=0)
        reference = None if lo == 0 else self.image2latent(images[0])
        lat = self.edit_latents(frame_latents, pos, neg, inv, num_inference_steps, strength, guidance_scale,
                                frame_batch=frame_batch, start_latents=None if latents is None else latents[lo:hi],
                                reference_latent=reference, use_cuda_graph=use_cuda_graph,
                                reference_start_latent=None if (latents is None or lo == 0) else latents[:1])

        if output_type == "latent":
            image = lat if shard is None else parallel.gather_frames(lat.contiguous(), num_frames)
        else:
            sf = self.vae.config.scaling_factor
            image = torch.cat([self.vae.decode(lat[i:i + 1] / sf).sample for i in range(hi - lo)], dim=0)      # :721-727
            if shard is not None:
                image = parallel.gather_frames(image.contiguous(), num_frames)
            if output_type != "pt":
                image = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy()
                if output_type == "pil":
                    from PIL import Image
                    image = [Image.fromarray((im * 255).round().astype("uint8")) for im in image]
        return StableDiffusionPipelineOutput(images=image) if return_dict else (image, None)
