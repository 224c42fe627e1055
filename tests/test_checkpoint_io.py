"""CPU: diffusers-layout checkpoint I/O (afldm_b200/hub.py) - the ``from_pretrained`` calls the reference's scripts
start with (scripts/shift_ldm_ffhq.py:165-170, shift_ldm_sr.py:156-161).  No kernels are launched."""
import json
import os

import pytest
import torch

from afldm_b200 import hub
from afldm_b200.af_modules import af_api
from afldm_b200.af_modules.af_blocks import AliasFreeDownsample2D, WarpedNonlinearity
from afldm_b200.models import AliasFreeAutoencoderKL, UNet2DModel
from afldm_b200.models.af_vae import AutoencoderKL
from afldm_b200.pipelines import MyLDMPipeline
from afldm_b200.pipelines.i2sb_pipeline import I2SBLDMPipeline
from afldm_b200.schedulers import DDIMScheduler
from afldm_b200.schedulers.i2sb import I2SBScheduler
from oracle import nn as ON

SMALL_UNET = dict(block_out_channels=[32, 64, 64], down_block_types=["AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
                  up_block_types=["UpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D"], attention_head_dim=8, sample_size=8)
SMALL_VAE = dict(block_out_channels=[32, 32, 64, 64])


def same_state(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k].cpu(), sb[k].cpu()), k


@pytest.mark.parametrize("safe", [True, False])
def test_unet_round_trip_and_diffusers_file_names(tmp_path, safe):
    torch.manual_seed(0)
    unet = UNet2DModel.from_config(SMALL_UNET)
    unet.save_pretrained(str(tmp_path / "unet"), safe_serialization=safe)
    names = sorted(os.listdir(tmp_path / "unet"))
    assert names == ["config.json", hub.WEIGHTS_SAFE if safe else hub.WEIGHTS_BIN]
    cfg = json.load(open(tmp_path / "unet" / "config.json"))
    assert cfg["_class_name"] == "UNet2DModel" and cfg["block_out_channels"] == [32, 64, 64]
    back = UNet2DModel.from_pretrained(str(tmp_path / "unet"))
    same_state(unet, back)
    assert not back.training
    # a checkpoint written by (the restatement of) diffusers loads strictly: same keys, same shapes
    ref = ON.UNet2DModel(**SMALL_UNET)
    from safetensors.torch import save_file
    os.makedirs(tmp_path / "d" / "unet")
    json.dump(dict(_class_name="UNet2DModel", _diffusers_version="0.32.1", **SMALL_UNET),
              open(tmp_path / "d" / "unet" / "config.json", "w"))
    save_file({k: v.contiguous() for k, v in ref.state_dict().items()}, str(tmp_path / "d" / "unet" / hub.WEIGHTS_SAFE))
    same_state(ref, UNet2DModel.from_pretrained(str(tmp_path / "d"), subfolder="unet"))


def test_pipeline_round_trip_then_surgery_like_the_reference_scripts(tmp_path):
    torch.manual_seed(1)
    pipe = MyLDMPipeline(AutoencoderKL.from_config(SMALL_VAE), UNet2DModel.from_config(SMALL_UNET), DDIMScheduler.from_config())
    pipe.save_pretrained(str(tmp_path / "p"))
    assert sorted(os.listdir(tmp_path / "p")) == ["model_index.json", "scheduler", "unet", "vae"]
    idx = json.load(open(tmp_path / "p" / "model_index.json"))
    assert idx["_class_name"] == "MyLDMPipeline" and idx["unet"] == ["diffusers", "UNet2DModel"]
    back = MyLDMPipeline.from_pretrained(str(tmp_path / "p"))
    same_state(pipe.unet, back.unet)
    same_state(pipe.vae, back.vae)
    assert dict(back.scheduler.config) == dict(pipe.scheduler.config)
    assert torch.equal(back.scheduler.alphas_cumprod, pipe.scheduler.alphas_cumprod)
    # plain modules come back (shift_ldm_ffhq.py:169-170 applies the surgery afterwards) ...
    assert not any(isinstance(m, WarpedNonlinearity) for m in back.unet.modules())
    af_api.make_af_unet(back.unet)
    af_api.make_af_vae_from_config(back.vae)
    assert any(isinstance(m, WarpedNonlinearity) for m in back.unet.modules())
    assert any(isinstance(m, AliasFreeDownsample2D) for m in back.vae.modules())
    same_state(pipe.unet, back.unet)                     # ... and the surgery leaves the state dict unchanged
    # the alias-free VAE class applies it in its constructor and reads the same files
    afvae = AliasFreeAutoencoderKL.from_pretrained(str(tmp_path / "p"), subfolder="vae")
    same_state(pipe.vae, afvae)
    assert afvae.config.scaling_factor == 0.6 and afvae.downsample_ratio == 8


def test_i2sb_pipeline_uses_its_scheduler(tmp_path):
    torch.manual_seed(2)
    pipe = I2SBLDMPipeline(None, UNet2DModel.from_config(SMALL_UNET), I2SBScheduler.from_config())
    pipe.save_pretrained(str(tmp_path / "sr"))
    back = I2SBLDMPipeline.from_pretrained(str(tmp_path / "sr"))
    assert isinstance(back.scheduler, I2SBScheduler) and back.vae is None
    assert torch.equal(back.scheduler.std_fwd, pipe.scheduler.std_fwd)


def test_repo_id_resolves_in_the_local_hub_cache_only(tmp_path, monkeypatch):
    snap = tmp_path / "hub" / "models--SingleZombie--alias_free_ldm_ffhq" / "snapshots" / "abc123"
    torch.manual_seed(3)
    pipe = MyLDMPipeline(None, UNet2DModel.from_config(SMALL_UNET), DDIMScheduler.from_config())
    pipe.save_pretrained(str(snap))
    monkeypatch.setenv("HF_HUB_CACHE", str(tmp_path / "hub"))
    assert hub.resolve("SingleZombie/alias_free_ldm_ffhq") == str(snap)
    back = MyLDMPipeline.from_pretrained("SingleZombie/alias_free_ldm_ffhq")        # the literal call of shift_ldm_ffhq.py:165
    same_state(pipe.unet, back.unet)
    with pytest.raises(FileNotFoundError, match="no network"):
        MyLDMPipeline.from_pretrained("SingleZombie/not_cached")


def test_strict_loading_rejects_a_mismatched_checkpoint(tmp_path):
    torch.manual_seed(4)
    unet = UNet2DModel.from_config(SMALL_UNET)
    unet.save_pretrained(str(tmp_path / "u"))
    cfg = json.load(open(tmp_path / "u" / "config.json"))
    cfg["block_out_channels"] = [32, 64, 128]
    json.dump(cfg, open(tmp_path / "u" / "config.json", "w"))
    with pytest.raises(RuntimeError):
        UNet2DModel.from_pretrained(str(tmp_path / "u"))
