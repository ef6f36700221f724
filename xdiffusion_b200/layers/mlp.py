"""Mlp parameter container + kernel forward (reference: layers/mlp.py:7-48)."""
import torch

from .. import ops
from .utils import Packed, bf16_weight


class Mlp(torch.nn.Module, Packed):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=None, bias=True, drop=0.0,
                 act="gelu_tanh", **kwargs):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = torch.nn.Linear(in_features, hidden_features, bias=bias)
        self.fc2 = torch.nn.Linear(hidden_features, out_features, bias=bias)
        self.act = {"gelu_tanh": ops.ACT_GELU, "silu": ops.ACT_SILU}[act]

    def weights(self):
        return self.packed("w", (self.fc1.weight, self.fc2.weight),
                           lambda: (bf16_weight(self.fc1.weight), bf16_weight(self.fc2.weight)))

    def forward(self, x_bf16, out_dtype=torch.bfloat16, ln=None, **epilogue):
        """x bf16 [M, in] -> fc2(act(fc1(x))); ``epilogue`` (gate/residual/out) applies to fc2.
        ``ln=(shift, scale, rows_per_mod)``: x is the fp32 residual stream and LayerNorm + modulate is fused into fc1."""
        w1, w2 = self.weights()
        if ln is not None:
            h = ops.ln_linear(x_bf16, ln[0], ln[1], ln[2], w1, self.fc1.bias, act=self.act)
        else:
            h = ops.linear(x_bf16, w1, self.fc1.bias, act=self.act)
        return ops.linear(h, w2, self.fc2.bias, out_dtype=out_dtype, **epilogue)
