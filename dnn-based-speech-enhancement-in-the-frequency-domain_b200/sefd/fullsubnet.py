"""Host side of the FullSubNet path: parameter containers with the reference's state_dict keys and RNG
consumption, and the autograd bridge onto sefd_fsn_forward / sefd_fsn_backward.

Reference being mirrored: models.py:568-682 (class FullSubNet), tools_for_model.py:726-795 (SequenceModel).
"""
import math

import torch
import torch.nn as nn

from . import _lib
from .ops import ptr, stream


class StackedLSTMParams(nn.Module):
    """Parameters of nn.LSTM(input_size, hidden_size, num_layers) under nn.LSTM's names, drawn U(-1/sqrt(H), 1/sqrt(H)) in
    nn.LSTM.reset_parameters' order (layer by layer: weight_ih, weight_hh, bias_ih, bias_hh)."""

    def __init__(self, input_size, hidden_size, num_layers=2, dropout=0.8):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers, self.dropout = input_size, hidden_size, num_layers, dropout
        k = 1.0 / math.sqrt(hidden_size)
        for l in range(num_layers):
            i = input_size if l == 0 else hidden_size
            for name, shape in ((f"weight_ih_l{l}", (4 * hidden_size, i)), (f"weight_hh_l{l}", (4 * hidden_size, hidden_size)),
                                (f"bias_ih_l{l}", (4 * hidden_size,)), (f"bias_hh_l{l}", (4 * hidden_size,))):
                setattr(self, name, nn.Parameter(torch.empty(*shape)))
        for p in self.parameters():
            nn.init.uniform_(p, -k, k)

    def flatten_parameters(self):
        pass


class SequenceModelParams(nn.Module):
    """tools_for_model.py:726-770: `sequence_model` (nn.LSTM, 2 layers, dropout 0.8) + `fc_output_layer` (nn.Linear)."""

    def __init__(self, input_size, output_size, hidden_size):
        super().__init__()
        from .dccrn import LinearParams
        self.sequence_model = StackedLSTMParams(input_size, hidden_size)
        self.fc_output_layer = LinearParams(hidden_size, output_size)


class _ForwardFSN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, noisy_mag, train, dropout_p, mask_fb, mask_sb, seed, *params):
        B, F, Tf = noisy_mag.shape
        plan = engine.plan(B, Tf)
        dev = noisy_mag.device
        ws = plan.workspace(dev)
        crm = torch.empty(B, F, Tf, 2, device=dev)
        plan.generation += 1
        _lib.check(_lib.load().sefd_fsn_forward(plan.handle, ptr(engine.flat), ptr(noisy_mag), int(train), float(dropout_p),
                                                ptr(mask_fb), ptr(mask_sb), int(seed), ptr(crm), ptr(ws), plan.ws_bytes, stream()),
                   "fsn_forward")
        ctx.engine, ctx.plan, ctx.generation = engine, plan, plan.generation
        ctx.masks = (mask_fb, mask_sb)          # injected masks are read again by the backward
        return crm

    @staticmethod
    def backward(ctx, g):
        engine, plan = ctx.engine, ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("sefd: the activation workspace of this forward was overwritten by a later forward "
                               "of the same batch shape; call backward() before the next forward")
        g = g.contiguous().float()
        _lib.check(_lib.load().sefd_fsn_backward(plan.handle, ptr(engine.flat), ptr(g), ptr(engine.flat_grad), ptr(plan.ws),
                                                 plan.ws_bytes, stream()), "fsn_backward")
        engine.backwards_since_step += 1
        flat = engine.flat_grad.clone()         # p.grad must not alias the buffer the next backward overwrites
        grads = tuple(flat[o: o + n].view(shape) for (_, o, n, shape) in plan.params)
        return (None,) * 7 + grads
