"""The module-swap ("surgery") API - the reference's drop-in boundary.

Public names, argument meaning and in-place / returns-None behaviour follow
/root/reference/afldm/af_modules/af_api.py:9-93: a model's ``block.downsamplers[0]`` /
``block.upsamplers[0]`` become alias-free blocks that keep the original ``conv`` Parameters, and
every resnet's single shared ``nonlinearity`` is wrapped in ``WarpedNonlinearity``.  What is
swapped in launches the sm_100a kernels (``afldm_b200.af_modules.af_blocks``).
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

from .af_blocks import AliasFreeDownsample2D, AliasFreeUpsample2D, WarpedNonlinearity

_ALL = (True,) * 8


def wrap_nonlinearity(nonlinearity):
    """af_api.py:9-10."""
    return WarpedNonlinearity(nonlinearity)


def replace_upsampler(ori_upsampler):
    """af_api.py:13-17: same channels / flags, re-using the trained ``conv``."""
    u = ori_upsampler
    return AliasFreeUpsample2D(u.channels, u.use_conv, out_channels=u.out_channels, ori_conv=u.conv)


def replace_downsampler(ori_downsampler):
    """af_api.py:20-26 (the conv's stride is forced to 1 inside the new block)."""
    d = ori_downsampler
    return AliasFreeDownsample2D(d.channels, d.use_conv, out_channels=d.out_channels, padding=d.padding,
                                 ori_conv=d.conv)


def wrap_resblock_nonlinearity(block):
    """af_api.py:29-31."""
    for resnet in block.resnets:
        resnet.nonlinearity = wrap_nonlinearity(resnet.nonlinearity)


def _operate(blocks: Iterable, attr: Optional[str], swap, resample_flags: Sequence[bool],
             act_flags: Sequence[bool]) -> None:
    """Walk ``blocks``: swap ``getattr(block, attr)[0]`` where flagged, wrap resnet activations where flagged."""
    for i, block in enumerate(blocks):
        samplers = getattr(block, attr, None) if attr else None
        if samplers is not None and resample_flags[i]:
            samplers[0] = swap(samplers[0])
        if act_flags[i]:
            wrap_resblock_nonlinearity(block)


def make_af_vae(vae, mod_mid_act=True, mod_down_filtered_act=(True, True, True, True),
                mod_up_filtered_act=(True, True, True, True), mod_resampling_layer=(True, True, True)):
    """af_api.py:34-60.  The encoder consumes the resampling flags in reverse order (:42); the last
    block of either side has no resampler, so a missing flag there is never read."""
    up_flags = list(mod_resampling_layer) + [False]
    down_flags = list(reversed(list(mod_resampling_layer))) + [False]
    _operate(vae.encoder.down_blocks, "downsamplers", replace_downsampler, down_flags, mod_down_filtered_act)
    if mod_mid_act:
        for side in (vae.encoder, vae.decoder):
            wrap_resblock_nonlinearity(side.mid_block)
    _operate(vae.decoder.up_blocks, "upsamplers", replace_upsampler, up_flags, mod_up_filtered_act)


def make_af_vae_from_config(vae):
    """af_api.py:63-67: flags come from the VAE's own config."""
    c = vae.config
    make_af_vae(vae, c.mid_act, c.down_filtered_act, c.up_filtered_act, c.up_rescale)


def make_af_unet(unet):
    """af_api.py:70-83: every resampler and every resnet activation of the UNet."""
    _operate(unet.down_blocks, "downsamplers", replace_downsampler, _ALL, _ALL)
    wrap_resblock_nonlinearity(unet.mid_block)
    _operate(unet.up_blocks, "upsamplers", replace_upsampler, _ALL, _ALL)


def make_af_controlnet(model):
    """af_api.py:86-93: the ControlNet has the UNet's down path and mid block only."""
    _operate(model.down_blocks, "downsamplers", replace_downsampler, _ALL, _ALL)
    wrap_resblock_nonlinearity(model.mid_block)
