"""Cross-frame attention: the state machine and attention processor of
/root/reference/afldm/pipelines/cross_frame_attn.py (AttnState :6-51, CrossFrameAttnProcessor
:54-130, processor get / set helpers :133-190) on top of the CUDA ``AttnProcessor2_0``.

STORE keeps each self-attention layer's (pre-GroupNorm) input per timestep; LOAD feeds the stored
reference-frame map, group-normalised, as the K/V source of later passes.  Where the reference
tiles the map over the batch (:91-97), the attention kernel here reads K/V batch ``b // (B/Bkv)``
directly, so nothing is materialised.
"""
from __future__ import annotations

from typing import Dict, Union

import torch

from .. import ops
from ..models.blocks import AttnProcessor2_0


class SlotContext:
    """Device-side addressing of the stored maps for captured steps: ``slot`` (int32 scalar on the device) is the index
    of the current step; every processor keeps its maps in one table [num_slots, ...] and moves them with
    ``ops.slot_copy`` (``afldm_slot_copy_f32``), so neither the STORE nor the LOAD pass reads a host value."""

    def __init__(self, num_slots: int, device):
        self.num_slots = int(num_slots)
        self.slot = torch.zeros(1, dtype=torch.int32, device=device)
        self._index = torch.arange(self.num_slots, dtype=torch.int32, device=device)

    def set_slot(self, i: int) -> None:
        self.slot.copy_(self._index[i:i + 1])             # device-to-device: legal between graph replays


class AttnState:
    STORE, LOAD, IDLE = 0, 1, 2

    def __init__(self):
        self.slots = None
        self.reset()

    def reset(self):
        self._state, self._timestep, self._store_id, self._alpha = AttnState.STORE, 0, 0, 0

    def enable_slots(self, num_slots: int, device) -> "SlotContext":
        """Switch the processors that share this state to step-indexed device tables (one slot per denoising step)
        instead of dictionaries keyed by ``t.item()`` - the form a captured CUDA graph can replay."""
        self.slots = SlotContext(num_slots, device)
        return self.slots

    def disable_slots(self) -> None:
        self.slots = None

    state = property(lambda self: self._state)
    timestep = property(lambda self: self._timestep)
    store_id = property(lambda self: self._store_id)
    alpha = property(lambda self: self._alpha)

    def set_timestep(self, t):
        self._timestep = t.item() if isinstance(t, torch.Tensor) else t

    def set_alpha(self, alpha):
        self._alpha = alpha

    def set_store_id(self, store_id):
        self._store_id = store_id

    def to_load(self):
        self._state = AttnState.LOAD

    def to_idle(self):
        self._state = AttnState.IDLE


class CrossFrameAttnProcessor(AttnProcessor2_0):
    def __init__(self, attn_state: AttnState, enable_interp: bool = False):
        self.attn_state = attn_state
        self.maps = [dict(), dict()]
        self.enable_interp = enable_interp
        self._tables = [None, None]        # slot mode: [num_slots, n, H, W, C] per store_id
        self._stage = [None, None]         # slot mode: the map of the current step, copied out of the table
        self._pairs = [{}, {}]             # slot mode: every (table, stage) ever allocated, by shape (graphs point at them)

    def _slot_store(self, store_id: int, x_nhwc: torch.Tensor) -> None:
        """One (table, stage) pair per stored-map SHAPE: a captured step keeps addressing the pair it was captured with
        when a later STORE pass has another batch (the inversion pass stores 1 map, the guided pass 2)."""
        ctx = self.attn_state.slots
        key = (ctx.num_slots, tuple(x_nhwc.shape), str(x_nhwc.device))
        pair = self._pairs[store_id].get(key)
        if pair is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("cross-frame map table would be allocated during CUDA-graph capture: warm up first")
            pair = self._pairs[store_id][key] = (
                torch.empty((ctx.num_slots,) + tuple(x_nhwc.shape), dtype=torch.float32, device=x_nhwc.device),
                torch.empty(tuple(x_nhwc.shape), dtype=torch.float32, device=x_nhwc.device))
        self._tables[store_id], self._stage[store_id] = pair          # what the next LOAD pass reads
        ops.slot_copy(pair[0], x_nhwc.contiguous(), ctx.slot, store=True)

    def _slot_load(self, store_id: int) -> torch.Tensor:
        tab, stage = self._tables[store_id], self._stage[store_id]
        if tab is None:
            raise RuntimeError("cross-frame attention: LOAD before any STORE pass")
        ops.slot_copy(tab, stage, self.attn_state.slots.slot, store=False)
        return stage.view(stage.shape[0], stage.shape[1] * stage.shape[2], stage.shape[3])      # [n, HW, C]

    @staticmethod
    def _kv_source(attn, stored: torch.Tensor) -> torch.Tensor:
        """Stored map ([n,C,H,W] or [n,HW,C]) -> group-normalised [n, HW, C] (cross_frame_attn.py:79-88)."""
        if stored.ndim == 4:
            x = ops.nhwc(stored)
            n, h, w, c = x.shape
            x = x.reshape(n, h * w, 1, c)
        else:
            x = stored.contiguous()
            x = x.view(x.shape[0], x.shape[1], 1, x.shape[2])
        if attn.group_norm is not None:
            gn = attn.group_norm
            x = ops.groupnorm_act(x, gn.num_groups, gn.eps, gn.weight, gn.bias, act="identity")
        return x.view(x.shape[0], x.shape[1], x.shape[3])

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 *args, **kwargs):
        st = self.attn_state
        plain = super().__call__
        if encoder_hidden_states is not None or st.state == AttnState.IDLE:
            return plain(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
        t = st.timestep
        slotted = st.slots is not None
        if st.state == AttnState.STORE:
            if slotted:
                hs = hidden_states
                self._slot_store(st.store_id, ops.nhwc(hs) if hs.ndim == 4 else
                                 hs.contiguous().view(hs.shape[0], hs.shape[1], 1, hs.shape[2]))
            else:
                self.maps[st.store_id][t] = hidden_states.detach()
            return plain(attn, hidden_states, None, attention_mask, temb)
        map0 = self._slot_load(0) if slotted else self.maps[0][t]
        res = plain(attn, hidden_states, self._kv_source(attn, map0), attention_mask, temb)
        if self.enable_interp:                                 # morphing between two stored frames (:100-122)
            map1 = self._slot_load(1) if slotted else self.maps[1][t]
            res2 = plain(attn, hidden_states, self._kv_source(attn, map1), attention_mask, temb)
            a = float(st.alpha)
            r1, r2 = ops.nhwc(res), ops.nhwc(res2)
            res = ops.nchw_view(ops.axpby(r1, r2, 1.0 - a, a))
        return res


def _named_attention_modules(model):
    for name, module in model.named_modules():
        if name and hasattr(module, "set_processor") and hasattr(module, "get_processor"):
            yield name, module


def get_unet_attn_processors(unet) -> Dict[str, object]:
    """{"<module path>.processor": processor} for every attention layer (cross_frame_attn.py:133-155)."""
    return {f"{name}.processor": m.get_processor() for name, m in _named_attention_modules(unet)}


def set_unet_attn_processor(unet, processor: Union[object, Dict[str, object]]) -> None:
    """Install one processor everywhere, or a dict keyed like ``get_unet_attn_processors`` (:158-190)."""
    mods = list(_named_attention_modules(unet))
    if isinstance(processor, dict):
        if len(processor) != len(mods):
            raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} "
                             f"does not match the number of attention layers: {len(mods)}.")
        for name, m in mods:
            m.set_processor(processor.pop(f"{name}.processor"))
    else:
        for _, m in mods:
            m.set_processor(processor)
