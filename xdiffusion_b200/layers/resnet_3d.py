"""Video residual block: (1,3,3) convolutions = per-frame 2-D convs over NHWC bf16 with frames folded
into the image count; GroupNorm statistics span all frames of a clip; the time embedding goes
through ``mlp_layers`` SiLU-Mlps (reference: layers/resnet_3d.py:103-254)."""
import torch

from .. import ops
from .mlp import Mlp
from .resnet import ResnetBlockBigGAN


class ResnetBlockBigGAN3D(ResnetBlockBigGAN):
    conv_dims = 3

    @staticmethod
    def _make_emb_layers(time_emb_dim, width, kwargs):
        n = kwargs.get("mlp_layers", 1)
        return torch.nn.Sequential(*[Mlp(in_features=time_emb_dim if i == 0 else width, out_features=width, act="silu")
                                     for i in range(n)])

    def emb_linear(self):
        raise NotImplementedError("3-D blocks run their embedding Mlp stack themselves")

    def embedding(self, temb_bf16):
        """temb bf16 [B, 4*nf] -> fp32 [B, 2*Cout] = [scale | shift]."""
        e = temb_bf16
        mlps = list(self.emb_layers)
        for i, m in enumerate(mlps):
            e = m(e, out_dtype=torch.float32 if i == len(mlps) - 1 else torch.bfloat16)
        return e
