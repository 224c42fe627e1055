"""diffusers-layout checkpoint I/O: ``from_pretrained`` / ``save_pretrained`` for the host mirror.

The reference's scripts start with ``MyLDMPipeline.from_pretrained('SingleZombie/alias_free_ldm_ffhq')``
(/root/reference/scripts/shift_ldm_ffhq.py:165-170; shift_ldm_sr.py:156-161) and then apply ``make_af_unet`` /
``make_af_vae_from_config`` to the loaded plain modules.  A diffusers pipeline directory is

    model_index.json
    unet/config.json        unet/diffusion_pytorch_model.safetensors   (or .bin)
    vae/config.json         vae/diffusion_pytorch_model.safetensors
    scheduler/scheduler_config.json

and the parameter names of ``afldm_b200.models`` are diffusers' own, so the tensors load with ``strict=True``.
There is no network here: a repo id is resolved in the local Hugging Face cache only
(``$HF_HUB_CACHE`` / ``$HF_HOME/hub`` / ``~/.cache/huggingface/hub``, ``models--<org>--<name>/snapshots/<rev>``).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch

WEIGHTS_SAFE = "diffusion_pytorch_model.safetensors"
WEIGHTS_BIN = "diffusion_pytorch_model.bin"
MODEL_CONFIG = "config.json"
SCHEDULER_CONFIG = "scheduler_config.json"
MODEL_INDEX = "model_index.json"
DIFFUSERS_VERSION = "0.32.1"          # the version recorded in the reference's shipped configs (configs/ldm/model_unet.json:3)


def resolve(path_or_repo: str) -> str:
    """A local directory, or the newest snapshot of ``org/name`` in the local Hugging Face cache."""
    if os.path.isdir(path_or_repo):
        return path_or_repo
    roots = [os.environ.get("HF_HUB_CACHE"),
             os.path.join(os.environ["HF_HOME"], "hub") if os.environ.get("HF_HOME") else None,
             os.path.expanduser("~/.cache/huggingface/hub")]
    folder = "models--" + path_or_repo.strip("/").replace("/", "--")
    for root in roots:
        if not root:
            continue
        snaps = os.path.join(root, folder, "snapshots")
        if os.path.isdir(snaps):
            revs = sorted((os.path.join(snaps, r) for r in os.listdir(snaps)), key=os.path.getmtime)
            if revs:
                return revs[-1]
    raise FileNotFoundError(
        f"'{path_or_repo}' is neither a directory nor a snapshot in the local Hugging Face cache "
        "(no network access here: download the checkpoint elsewhere and pass its path)")


def read_config(directory: str, name: str = MODEL_CONFIG) -> dict:
    with open(os.path.join(directory, name)) as f:
        cfg = json.load(f)
    return {k: v for k, v in cfg.items() if not k.startswith("_")}


def load_state_dict(directory: str) -> dict:
    safe, binf = os.path.join(directory, WEIGHTS_SAFE), os.path.join(directory, WEIGHTS_BIN)
    if os.path.exists(safe):
        from safetensors.torch import load_file
        return load_file(safe)
    if os.path.exists(binf):
        return torch.load(binf, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no {WEIGHTS_SAFE} / {WEIGHTS_BIN} in {directory}")


def load_model(cls, path_or_repo: str, subfolder: Optional[str] = None, strict: bool = True):
    """``cls.from_config(config.json)`` + ``load_state_dict`` (fp32; strict by default: the key sets are diffusers')."""
    d = resolve(path_or_repo)
    if subfolder:
        d = os.path.join(d, subfolder)
    model = cls.from_config(read_config(d))
    sd = {k: v.to(torch.float32) for k, v in load_state_dict(d).items()}
    model.load_state_dict(sd, strict=strict)
    return model.eval()


def save_model(model, directory: str, class_name: str, safe_serialization: bool = True) -> None:
    os.makedirs(directory, exist_ok=True)
    cfg = {"_class_name": class_name, "_diffusers_version": DIFFUSERS_VERSION}
    cfg.update({k: v for k, v in dict(model.config).items()})
    with open(os.path.join(directory, MODEL_CONFIG), "w") as f:
        json.dump(cfg, f, indent=2)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in model.state_dict().items()}
    if safe_serialization:
        from safetensors.torch import save_file
        save_file(sd, os.path.join(directory, WEIGHTS_SAFE))
    else:
        torch.save(sd, os.path.join(directory, WEIGHTS_BIN))


def load_scheduler(cls, path_or_repo: str, subfolder: Optional[str] = "scheduler"):
    d = resolve(path_or_repo)
    if subfolder and os.path.isdir(os.path.join(d, subfolder)):
        d = os.path.join(d, subfolder)
    return cls.from_config(read_config(d, SCHEDULER_CONFIG))


def save_scheduler(scheduler, directory: str, class_name: str) -> None:
    os.makedirs(directory, exist_ok=True)
    cfg = {"_class_name": class_name, "_diffusers_version": DIFFUSERS_VERSION}
    cfg.update(dict(scheduler.config))
    with open(os.path.join(directory, SCHEDULER_CONFIG), "w") as f:
        json.dump(cfg, f, indent=2)
