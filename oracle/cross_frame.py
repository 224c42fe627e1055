"""Oracle (test infrastructure): cross-frame attention state machine.

Follows /root/reference/afldm/pipelines/cross_frame_attn.py:6-51 (AttnState) and :54-130
(CrossFrameAttnProcessor) on top of ``oracle.nn.AttnProcessor2_0``.
"""
from __future__ import annotations

import torch

from .nn import AttnProcessor2_0


class AttnState:
    STORE, LOAD, IDLE = 0, 1, 2

    def __init__(self):
        self.reset()

    def reset(self):
        self.state, self.timestep, self.store_id, self.alpha = AttnState.STORE, 0, 0, 0

    def set_timestep(self, t):
        self.timestep = t.item() if isinstance(t, torch.Tensor) else t

    def set_alpha(self, a):
        self.alpha = a

    def set_store_id(self, i):
        self.store_id = i

    def to_load(self):
        self.state = AttnState.LOAD

    def to_idle(self):
        self.state = AttnState.IDLE


class CrossFrameAttnProcessor(AttnProcessor2_0):
    def __init__(self, attn_state: AttnState, enable_interp=False, base=None):
        """``base``: the plain processor this one extends (default: the 4-D attention-block form of oracle.nn; pass
        ``oracle.nn_cond.AttnProcessor2_0()`` for the 3-D transformer form of the SD-1.5 UNet)."""
        self.attn_state = attn_state
        self.maps = [dict(), dict()]
        self.enable_interp = enable_interp
        self.base = base

    def _kv_source(self, attn, m, batch):
        # cross_frame_attn.py:79-97: (n,c,h,w) -> (n,hw,c), group-norm it, tile over the batch
        if m.ndim == 4:
            n, c, h, w = m.shape
            m = m.view(n, c, h * w).transpose(1, 2)
        if attn.group_norm is not None:
            m = attn.group_norm(m.transpose(1, 2)).transpose(1, 2)
        if m.shape[0] < batch:
            n, s, c = m.shape
            m = m.unsqueeze(1).repeat(1, batch // n, 1, 1).reshape(batch, s, c)
        return m

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        st = self.attn_state
        base = super().__call__ if self.base is None else self.base
        if encoder_hidden_states is not None or st.state == AttnState.IDLE:
            return base(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
        t = st.timestep
        if st.state == AttnState.STORE:
            self.maps[st.store_id][t] = hidden_states.detach()
            return base(attn, hidden_states, None, attention_mask, temb)
        b = hidden_states.shape[0]
        out0 = base(attn, hidden_states, self._kv_source(attn, self.maps[0][t], b), attention_mask, temb)
        if not self.enable_interp:
            return out0
        out1 = base(attn, hidden_states, self._kv_source(attn, self.maps[1][t], b), attention_mask, temb)
        return (1 - st.alpha) * out0 + st.alpha * out1


def set_attn_processor(model, make):
    """Install ``make()`` on every module that exposes ``set_processor`` (cross_frame_attn.py:158-190)."""
    n = 0
    for m in model.modules():
        if hasattr(m, "set_processor"):
            m.set_processor(make())
            n += 1
    return n
