"""UNet2DModel with the diffusers surface (``unet(sample, t).sample``, ``.config``, ``.dtype``,
``down_blocks / mid_block / up_blocks`` layout that ``make_af_unet`` walks) whose forward is the
sm_100a kernel sequence of one denoising step (reference call site
afldm/pipelines/ldm_pipeline.py:103-109, scripts/shift_ldm_ffhq.py:98)."""
from __future__ import annotations

from typing import Sequence, Union

import torch
import torch.nn as nn

from .. import ops
from ..af_modules.af_blocks import act_name
from ..configs import Config, FFHQ_UNET
from ..packing import conv_params, conv_params_padded, fused_linear_params
from .blocks import DownBlock2D, UNetMidBlock2D, UpBlock2D


class UNet2DOutput:
    def __init__(self, sample):
        self.sample = sample


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample: torch.Tensor) -> torch.Tensor:
        h = ops.linear_rows(sample, self.linear_1.weight, self.linear_1.bias)
        return ops.linear_rows(h, self.linear_2.weight, self.linear_2.bias, act_in="silu")


_SIDE_STREAMS = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    """One auxiliary stream per device for work that is independent of the activation path."""
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)      # one per calling stream: parallel branches stay parallel
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=idx)
    return st


class UNet2DModel(nn.Module):
    def __init__(self, sample_size=32, in_channels=4, out_channels=4,
                 block_out_channels: Sequence[int] = (192, 384, 384, 768, 768),
                 down_block_types=("AttnDownBlock2D",) * 4 + ("DownBlock2D",),
                 up_block_types=("UpBlock2D",) + ("AttnUpBlock2D",) * 4,
                 layers_per_block=2, attention_head_dim=24, norm_eps=1e-5, norm_num_groups=32,
                 downsample_padding=1, flip_sin_to_cos=True, freq_shift=0, act_fn="silu", **unused):
        super().__init__()
        if not flip_sin_to_cos or freq_shift != 0 or act_fn != "silu" or norm_num_groups != 32:
            raise NotImplementedError("only the AF-LDM UNet configuration family is supported")
        boc = list(block_out_channels)
        self.config = Config(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                             block_out_channels=boc, down_block_types=list(down_block_types),
                             up_block_types=list(up_block_types), layers_per_block=layers_per_block,
                             attention_head_dim=attention_head_dim, norm_eps=norm_eps,
                             norm_num_groups=norm_num_groups, downsample_padding=downsample_padding,
                             flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift, act_fn=act_fn)
        temb_ch = boc[0] * 4
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, kind in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownBlock2D(
                in_ch, out_ch, temb_ch, layers_per_block, add_downsample=i != len(boc) - 1,
                attention_head_dim=attention_head_dim if kind.startswith("Attn") else None,
                eps=norm_eps, downsample_padding=downsample_padding))
        self.mid_block = UNetMidBlock2D(boc[-1], temb_ch, attention_head_dim, norm_eps)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_ch = rev[0]
        for i, kind in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock2D(
                in_ch, out_ch, prev, temb_ch, layers_per_block + 1, add_upsample=i != len(boc) - 1,
                attention_head_dim=attention_head_dim if kind.startswith("Attn") else None, eps=norm_eps))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @classmethod
    def from_config(cls, config=FFHQ_UNET, **overrides):
        cfg = dict(config)
        cfg.update(overrides)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    @classmethod
    def from_pretrained(cls, path_or_repo: str, subfolder=None, **_unused):
        """diffusers ``UNet2DModel.from_pretrained``: config.json + safetensors of a local directory / cached repo
        (plain module; the caller applies ``make_af_unet`` as the reference's scripts do)."""
        from .. import hub
        return hub.load_model(cls, path_or_repo, subfolder)

    def save_pretrained(self, directory: str, safe_serialization: bool = True):
        from .. import hub
        hub.save_model(self, directory, "UNet2DModel", safe_serialization)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def _resnets(self):
        for blk in self.down_blocks:
            yield from blk.resnets
        yield from self.mid_block.resnets
        for blk in self.up_blocks:
            yield from blk.resnets

    def _time_projections(self, emb: torch.Tensor):
        """time_emb_proj(silu(emb)) of every resnet in ONE launch; returns per-resnet [B, Cout] views."""
        resnets = list(self._resnets())
        w, b = fused_linear_params(self, "temb_proj", [r.time_emb_proj for r in resnets])
        allp = ops.linear_rows(emb, w, b, act_in="silu")
        out, off = [], 0
        for r in resnets:
            out.append(allp[:, off:off + r.out_channels])
            off += r.out_channels
        return out

    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int], return_dict: bool = True):
        bsz = sample.shape[0]
        dev = sample.device
        if torch.is_tensor(timestep):
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
            if t.numel() == 1 and bsz > 1:
                t = t.expand(bsz)
            t = t.contiguous()
        else:
            t = torch.full((bsz,), float(timestep), dtype=torch.float32, device=dev)
        # The time-embedding MLP and the 27 fused time_emb_proj rows depend on t only: they run on a side stream
        # (a parallel branch of the captured step graph) under conv_in and the first GroupNorm / filtered activation,
        # and are joined just before the first resnet consumes its projection row.
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            emb = self.time_embedding(ops.timestep_embedding(t, self.config.block_out_channels[0]))
            projs = self._time_projections(emb)

        x = ops.nhwc(sample)
        cin = x.shape[-1]
        if ops.default_conv_algo() == "tf32" and cin % 32 != 0:
            # 4 latent channels -> one zero-padded 32-channel chunk: conv_in on the tensor-core path, which also
            # emits the GroupNorm partial sums of its output (first resnet's norm1, last up block's skip concat)
            w, b, k = conv_params_padded(self.conv_in, 32 * ((cin + 31) // 32))
            h = ops.nchw_view(ops.conv2d(ops.pad_channels(x, w.shape[-1]), w, b, k, gn_stats=True))
        else:
            w, b, k = conv_params(self.conv_in)
            h = ops.nchw_view(ops.conv2d(x, w, b, k))
        ops.defer_join(main, side)      # joined by the first op that reads a projection row (first resnet's conv1)
        skips = (h,)
        pi = 0
        for blk in self.down_blocks:
            n = len(blk.resnets)
            h, outs = blk(h, emb, projs[pi:pi + n])
            pi += n
            skips += outs
        h = self.mid_block(h, emb, projs[pi:pi + 2])
        pi += 2
        for blk in self.up_blocks:
            n = len(blk.resnets)
            h = blk(h, skips[-n:], emb, projs[pi:pi + n])
            pi += n
            skips = skips[:-n]

        ops.join_deferred()             # no-op unless no resnet consumed a projection row
        hx = ops.nhwc(h)
        gn = self.conv_norm_out
        a = ops.groupnorm_act(hx, gn.num_groups, gn.eps, gn.weight, gn.bias, act=act_name(self.conv_act))   # conv_act is NOT wrapped
        w, b, k = conv_params(self.conv_out)
        out = ops.nchw_view(ops.conv2d(a, w, b, k))
        return UNet2DOutput(out) if return_dict else (out,)
