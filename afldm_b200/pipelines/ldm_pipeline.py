"""Unconditional latent-diffusion pipeline with the call surface of
/root/reference/afldm/pipelines/ldm_pipeline.py (``MyLDMPipeline.__call__`` :32-131,
``ddim_inversion`` :133-160; driven by scripts/shift_ldm_ffhq.py:50-159).

The per-step work (UNet forward + DDIM update) is a fixed sequence of sm_100a kernel launches,
so it is captured ONCE in a CUDA graph (``GraphedDenoiser``) and replayed for every timestep:
the timestep and the two DDIM coefficients are read from small device tensors that are refreshed
between replays, which removes the ~600 launches' worth of host overhead from every step and the
``t.item()`` synchronisation the reference pays (cross_frame_attn.py:31-33).
"""
from __future__ import annotations

from typing import List, Optional, Union

import os

import torch

from .. import ops
from ..af_modules.af_api import make_af_unet
from ..configs import FFHQ_AFVAE, FFHQ_DDIM, FFHQ_UNET
from ..models.af_vae import AliasFreeAutoencoderKL
from ..models.unet_2d import UNet2DModel
from ..schedulers.ddim import DDIMScheduler


class ImagePipelineOutput:
    def __init__(self, images):
        self.images = images


def randn_tensor(shape, generator=None, device=None, dtype=torch.float32):
    """diffusers.utils.torch_utils.randn_tensor: sample on the generator's device (CPU by default) so
    that seeds give the same latents everywhere, then move (ldm_pipeline.py:82-87)."""
    gdev = generator.device if generator is not None else torch.device("cpu")
    x = torch.randn(shape, generator=generator, device=gdev, dtype=dtype)
    return x.to(device) if device is not None else x


def graph_signature(unet) -> tuple:
    """Everything a captured step bakes in: the storage and version of every parameter (the packed weight copies of
    ``packing`` are rebuilt from them, and the capture holds raw pointers to those copies), the attention processors,
    the convolution algorithm and the feature switches.  ``MyLDMPipeline.graphed`` re-captures when it changes."""
    params = tuple((p.data_ptr(), p._version) for p in unet.parameters())
    procs = []
    for m in unet.modules():
        if hasattr(m, "get_processor"):
            pr = m.get_processor()
            st = getattr(pr, "attn_state", None)
            procs.append((id(pr), None if st is None else (st.state, st.store_id, id(st.slots),
                                                           bool(getattr(pr, "enable_interp", False)))))
    procs = tuple(procs)
    return (params, procs, ops.default_conv_algo(), ops.F16_CONV, ops.F16_ATTENTION, ops.FUSE_GN_PROLOGUE,
            ops.FUSE_CONCAT, ops.SHORTCUT_SIDE_STREAM)


_UNSET = object()


def _cfa_mode(unet):
    """(state, store_id) of the cross-frame processors installed on the UNet, None for the default processors."""
    for m in unet.modules():
        if hasattr(m, "get_processor"):
            st = getattr(m.get_processor(), "attn_state", None)
            if st is not None:
                return (st.state, st.store_id)
    return None


def graph_capturable(unet) -> bool:
    """The captured step covers the default attention processors, and ``CrossFrameAttnProcessor`` in slot mode
    (``AttnState.enable_slots``): in its reference form the processor keys host-side dictionaries by ``t.item()`` and (in
    STORE state) would keep pointers into the graph's private memory pool."""
    from ..models.blocks import AttnProcessor2_0
    from .cross_frame_attn import AttnState, CrossFrameAttnProcessor
    for m in unet.modules():
        if not hasattr(m, "get_processor"):
            continue
        pr = m.get_processor()
        if type(pr) is AttnProcessor2_0:
            continue
        # a cross-frame processor is capturable in slot mode (device-indexed map tables), with a fixed interpolation
        # weight of 0 / no interpolation (alpha is a host scalar baked into the capture)
        if type(pr) is CrossFrameAttnProcessor and pr.attn_state.slots is not None and not pr.enable_interp \
                and pr.attn_state.state in (AttnState.STORE, AttnState.LOAD):
            continue
        return False
    return True


class GraphedDenoiser:
    """One denoising step ``x <- cx * x + ce * unet(x, t)`` captured in a CUDA graph.

    ``x`` (NHWC, updated in place), ``t`` [B] and ``coef`` [2] are static device buffers."""

    def __init__(self, unet: UNet2DModel, batch: int, warmup: int = 2, size: Optional[int] = None):
        dev = unet.device
        self.signature = graph_signature(unet)
        c, s = unet.config.in_channels, (size or unet.config.sample_size)
        self.unet = unet
        self.x = torch.zeros((batch, s, s, c), dtype=torch.float32, device=dev)
        self.t = torch.ones((batch,), dtype=torch.float32, device=dev)
        self.coef = torch.tensor([1.0, 0.0], dtype=torch.float32, device=dev)
        self.eps = None
        self.branches = 2 if (os.environ.get("AFLDM_BRANCHES", "1") == "2" and batch % 2 == 0 and batch >= 4) else 1
        self._fork = torch.cuda.Stream(device=dev)
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

    def _body(self):
        if self.branches == 2:
            # two half batches as two parallel branches of the graph: the launch-latency-bound kernels of the
            # 8x8 ... 2x2 levels (few CTAs each) and the inter-kernel gaps of one branch overlap the other's work
            h = self.x.shape[0] // 2
            main = torch.cuda.current_stream(self.x.device)
            self._fork.wait_stream(main)
            with torch.cuda.stream(self._fork), ops.scratch_slot(1):
                self._half(h, self.x.shape[0])
            self._half(0, h)
            main.wait_stream(self._fork)
            return
        self._half(0, self.x.shape[0])

    def _half(self, lo: int, hi: int):
        x = self.x[lo:hi]
        eps = self.unet(ops.nchw_view(x), self.t[lo:hi], return_dict=False)[0]
        ops.axpby(x, ops.nhwc(eps), self.coef, None, out=x)

    def replay(self):
        self.graph.replay()


class MyLDMPipeline:
    def __init__(self, vae: AliasFreeAutoencoderKL, unet: UNet2DModel, scheduler: DDIMScheduler):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self._graphs = {}
        self._tables = {}
        self._graph_ok = None
        self._traj_state = None
        self._cfa_cached = _UNSET
        self._sweep = {}
        self._bar = {}

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_config(cls, unet_config=FFHQ_UNET, vae_config=FFHQ_AFVAE, scheduler_config=FFHQ_DDIM,
                    seed: int = 0, alias_free: bool = True, with_vae: bool = True):
        """Randomly initialised pipeline of the released architecture (no checkpoint is reachable
        offline).  Weights follow PyTorch's default initialisers under ``seed``."""
        torch.manual_seed(seed)
        unet = UNet2DModel.from_config(unet_config)
        if alias_free:
            make_af_unet(unet)
        vae = AliasFreeAutoencoderKL.from_config(vae_config) if with_vae else None
        return cls(vae, unet, DDIMScheduler.from_config(scheduler_config))

    SCHEDULER_CLS = DDIMScheduler

    @classmethod
    def from_pretrained(cls, path_or_repo: str, **_unused):
        """``DiffusionPipeline.from_pretrained`` for the diffusers directory layout (unet/, vae/, scheduler/; local path or a
        repo id present in the local Hugging Face cache).  Like the reference, this returns PLAIN modules: the scripts
        then call ``make_af_unet(pipe.unet)`` and ``make_af_vae_from_config(pipe.vae)`` (shift_ldm_ffhq.py:165-170)."""
        import os
        from .. import hub
        from ..models.af_vae import AutoencoderKL
        d = hub.resolve(path_or_repo)
        unet = UNet2DModel.from_pretrained(d, subfolder="unet")
        vae = AutoencoderKL.from_pretrained(d, subfolder="vae") if os.path.isdir(os.path.join(d, "vae")) else None
        return cls(vae, unet, cls.SCHEDULER_CLS.from_pretrained(d, subfolder="scheduler"))

    def save_pretrained(self, directory: str, safe_serialization: bool = True):
        import json
        import os
        os.makedirs(directory, exist_ok=True)
        index = {"_class_name": type(self).__name__, "_diffusers_version": "0.32.1",
                 "unet": ["diffusers", "UNet2DModel"], "scheduler": ["diffusers", type(self.scheduler).__name__]}
        self.unet.save_pretrained(os.path.join(directory, "unet"), safe_serialization)
        if self.vae is not None:
            index["vae"] = ["diffusers", "AutoencoderKL"]
            self.vae.save_pretrained(os.path.join(directory, "vae"), safe_serialization)
        self.scheduler.save_pretrained(os.path.join(directory, "scheduler"))
        with open(os.path.join(directory, "model_index.json"), "w") as f:
            json.dump(index, f, indent=2)

    def to(self, device):
        self.unet.to(device)
        if self.vae is not None:
            self.vae.to(device)
        self._graphs.clear()
        self._tables = {}
        self._sweep = {}
        return self

    @property
    def device(self):
        return self.unet.device

    def set_progress_bar_config(self, **kwargs):
        self._bar = kwargs

    def progress_bar(self, iterable):
        if self._bar.get("disable", True):
            return iterable
        from tqdm import tqdm
        return tqdm(iterable, **{k: v for k, v in self._bar.items() if k != "disable"})

    # ------------------------------------------------------------------ denoising
    def graphed(self, batch: int, check: bool = True, size: Optional[int] = None) -> GraphedDenoiser:
        """The captured step for this batch size; re-captured when anything it baked in has changed since
        (``load_state_dict``, an in-place weight update, other attention processors, another conv algorithm).
        The check walks the module tree (~2 ms of host time): ``denoise`` runs it at the start of a trajectory
        (``start == 0``), not for every window of one."""
        key = batch if size in (None, self.unet.config.sample_size) else (batch, size)
        if check or self._cfa_cached is _UNSET:
            self._cfa_cached = _cfa_mode(self.unet)
        mode = self._cfa_cached
        if mode is not None:
            key = (key, mode)                      # STORE / LOAD passes of the cross-frame processors are different graphs
        g = self._graphs.get(key)
        if g is not None and check and g.signature != graph_signature(self.unet):
            g = None
        if g is None:
            g = self._graphs[key] = GraphedDenoiser(self.unet, batch, size=size)
        return g

    def step_tables(self, num_inference_steps: int, batch: int):
        """Device tables [steps, B] of timesteps and [steps, 2] of DDIM coefficients."""
        self.scheduler.set_timesteps(num_inference_steps)
        ts = self.scheduler.timesteps
        coefs = torch.stack([torch.stack(self.scheduler.coefficients(int(t))) for t in ts]).to(torch.float32)
        tt = ts.to(torch.float32)[:, None].expand(-1, batch).contiguous()
        return tt.to(self.device), coefs.to(self.device)

    @torch.no_grad()
    def denoise(self, latents: torch.Tensor, num_inference_steps: int = 50, use_cuda_graph: bool = True,
                start: int = 0, stop: Optional[int] = None):
        """DDIM loop of ldm_pipeline.py:103-109 (eta = 0) on latents [B,C,H,W] (device or - pinned - host memory; the
        result stays on the device).  ``start`` / ``stop`` select a window [start, stop) of the schedule's steps, so a
        trajectory can be advanced piecewise.  The captured-graph path needs the default attention processors; with
        any other processor installed the loop runs eagerly."""
        latents = latents.to(device=self.device, dtype=torch.float32, non_blocking=True)
        stop = num_inference_steps if stop is None else stop
        if start == 0 or self._graph_ok is None:        # decided at the start of a trajectory, kept for its windows
            self._graph_ok = graph_capturable(self.unet)     # (these walk the module tree: ~1 ms of host time)
            self._traj_state = self._attn_state()
        if not use_cuda_graph or not self._graph_ok:
            self.scheduler.set_timesteps(num_inference_steps)
            state = self._traj_state
            slots = None if state is None else state.slots
            for i, t in zip(range(start, stop), self.progress_bar(self.scheduler.timesteps[start:stop])):
                if slots is not None:
                    slots.set_slot(i)
                elif state is not None:
                    state.set_timestep(int(t))          # dictionary mode keys the maps by timestep (shift_ldm_ffhq.py:96)
                eps = self.unet(self.scheduler.scale_model_input(latents, t), int(t)).sample
                latents = self.scheduler.step(eps, int(t), latents).prev_sample
            return ops.to_nchw_contiguous(ops.nhwc(latents))
        g = self.graphed(latents.shape[0], check=start == 0, size=latents.shape[-1])
        key = (num_inference_steps, latents.shape[0])
        if self._tables.get("key") != key:
            self._tables = {"key": key, "tt": None}
            self._tables["tt"], self._tables["coef"] = self.step_tables(num_inference_steps, latents.shape[0])
        tt, coefs = self._tables["tt"], self._tables["coef"]
        slots = None if self._traj_state is None else self._traj_state.slots
        g.x.copy_(ops.nhwc(latents))
        for i in self.progress_bar(range(start, stop)):
            g.t.copy_(tt[i])
            g.coef.copy_(coefs[i])
            if slots is not None:
                slots.set_slot(i)
            g.replay()
        return ops.to_nchw_contiguous(g.x)

    def _attn_state(self):
        for m in self.unet.modules():
            if hasattr(m, "get_processor"):
                st = getattr(m.get_processor(), "attn_state", None)
                if st is not None:
                    return st
        return None

    def _slot_context(self):
        st = self._attn_state()
        return None if st is None else st.slots

    @torch.no_grad()
    def shift_sweep(self, init_latent: torch.Tensor, shifts, num_inference_steps: int = 50, latent_shifter=None,
                    use_cuda_graph: bool = True):
        """The measurement loop of scripts/shift_ldm_ffhq.py:50-159 as TWO trajectories instead of 1 + len(shifts)
        sequential ones: the reference pass (cross-frame attention in STORE state, batch 1) and ONE batched pass over
        every shifted start latent (LOAD state; the stored maps of the single reference frame serve all of them, K / V
        batch 1).  Both replay captured steps: the per-timestep maps live in device tables indexed by a device-side
        step counter (``AttnState.enable_slots``), so there is no ``t.item()`` synchronisation and no host dictionary.

        ``init_latent`` [1,C,H,W]; ``shifts`` a sequence of (ti, tj) in latent pixels; ``latent_shifter`` defaults to
        ``ImageShifter('ideal_crop', 8)`` (:62).  Returns (denoised reference [1,C,H,W], denoised shifted [S,C,H,W],
        shifted start latents [S,C,H,W], masks [S,C,H,W]); the original attention processors are restored (:153-157)."""
        from ..shift_utils import ImageShifter
        from .cross_frame_attn import (AttnState, CrossFrameAttnProcessor, get_unet_attn_processors,
                                       set_unet_attn_processor)
        if init_latent.shape[0] != 1:
            raise ValueError("shift_sweep: one reference latent [1,C,H,W] expected")
        dev = self.device
        init_latent = init_latent.to(device=dev, dtype=torch.float32)
        shifter = latent_shifter or ImageShifter("ideal_crop", 8)
        prev = get_unet_attn_processors(self.unet)
        cache = self._sweep.get(num_inference_steps)
        if cache is None or cache[2] != tuple(prev.keys()):
            state = AttnState()
            state.enable_slots(num_inference_steps, dev)
            cache = self._sweep[num_inference_steps] = (state, {k: CrossFrameAttnProcessor(state) for k in prev},
                                                        tuple(prev.keys()))
        state, procs, _ = cache
        set_unet_attn_processor(self.unet, dict(procs))
        try:
            state.reset()                                           # STORE
            base = self.denoise(init_latent, num_inference_steps, use_cuda_graph)
            state.to_load()
            shifted, masks = shifter.shift_batch(init_latent, list(shifts))          # [S,1,C,H,W]
            shifted, masks = shifted[:, 0].contiguous(), masks[:, 0].contiguous()
            outs = self.denoise(shifted, num_inference_steps, use_cuda_graph)
        finally:
            set_unet_attn_processor(self.unet, dict(prev))
        return base, outs, shifted, masks

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
                 eta: float = 0.0, num_inference_steps: int = 50, latents=None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, use_cuda_graph: bool = True, **kwargs):
        if eta != 0.0:
            raise NotImplementedError("eta > 0 is not on the AF-LDM path")
        if latents is None:
            latents = randn_tensor((batch_size, self.unet.config.in_channels, self.unet.config.sample_size,
                                    self.unet.config.sample_size), generator=generator)
        latents = latents.to(device=self.device, dtype=self.unet.dtype) * self.scheduler.init_noise_sigma
        latents = self.denoise(latents, num_inference_steps, use_cuda_graph)
        if output_type == "latent":
            return latents
        image = self.decode_latents(latents)
        if output_type != "pt":
            image = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy()
            if output_type == "pil":
                from PIL import Image
                image = [Image.fromarray((im * 255).round().astype("uint8")) for im in image]
            return ImagePipelineOutput(images=image) if return_dict else (image,)
        return image

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """ldm_pipeline.py:117-119."""
        return self.vae.decode(latents / self.vae.config.scaling_factor).sample

    @torch.no_grad()
    def ddim_inversion(self, latent: torch.Tensor, bar: bool = True, use_cuda_graph: bool = False) -> torch.Tensor:
        """ldm_pipeline.py:133-160: deterministic DDIM run backwards over ``scheduler.timesteps``.  The update has the
        form of the sampling step, latent <- cx latent + ce eps(latent, t), so ``use_cuda_graph=True`` replays the
        captured step of ``denoise`` with the inversion's coefficients (default processors or slot-mode CFA only)."""
        sch = self.scheduler
        timesteps = sch.timesteps.flip(0)
        latent = latent.to(device=self.device, dtype=torch.float32)
        coefs = []
        for i, t in enumerate(timesteps):
            a_t = sch.alphas_cumprod[int(t)]
            a_prev = sch.alphas_cumprod[int(timesteps[i - 1])] if i > 0 else sch.final_alpha_cumprod
            mu, mu_prev = a_t ** 0.5, a_prev ** 0.5
            sigma, sigma_prev = (1 - a_t) ** 0.5, (1 - a_prev) ** 0.5
            # latent' = mu * (latent - sigma_prev * eps) / mu_prev + sigma * eps
            coefs.append((float(mu / mu_prev), float(sigma - mu * sigma_prev / mu_prev)))
        if use_cuda_graph and graph_capturable(self.unet) and self._attn_state() is None:
            g = self.graphed(latent.shape[0], size=latent.shape[-1])
            table = torch.tensor(coefs, dtype=torch.float32, device=self.device)
            tt = timesteps.to(device=self.device, dtype=torch.float32)[:, None].expand(-1, latent.shape[0]).contiguous()
            g.x.copy_(ops.nhwc(latent))
            for i in range(len(coefs)):
                g.t.copy_(tt[i])
                g.coef.copy_(table[i])
                g.replay()
            return ops.to_nchw_contiguous(g.x)
        for (cx, ce), t in zip(coefs, timesteps):
            eps = self.unet(latent, int(t)).sample
            latent = ops.nchw_view(ops.axpby(ops.nhwc(latent), ops.nhwc(eps), cx, ce))
        return ops.to_nchw_contiguous(ops.nhwc(latent))
