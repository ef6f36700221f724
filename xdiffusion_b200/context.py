"""Context adapters on the sampling path (reference: xdiffusion/context.py:40-61,72-104,160-177)."""
from typing import Dict, List

import torch


class NullContextAdapter(torch.nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, context: Dict):
        return None


class IgnoreContextAdapter(torch.nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, context: Dict, *args, **kwargs):
        return context


class IgnoreInputPreprocessor(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x, *args, **kwargs):
        return x


class UnconditionalClassesAdapter(torch.nn.Module):
    """classes -> the null class index ``num_classes`` for every sample."""

    def __init__(self, num_classes, **kwargs):
        super().__init__()
        self._num_classes = num_classes

    def forward(self, context: Dict, **kwargs):
        new_context = context.copy()
        new_context["classes"] = torch.zeros_like(context["classes"]) + self._num_classes
        return new_context


class UnconditionalTextPromptsAdapter(torch.nn.Module):
    def forward(self, context: Dict):
        new_context = context.copy()
        new_context["text_prompts"] = [""] * len(context["text_prompts"])
        return new_context


class UnconditionalEmbeddingAdapter(torch.nn.Module):
    """Learned null embedding (num_tokens, C) tiled over the batch."""

    def __init__(self, embedding_shape: List[int]):
        super().__init__()
        assert len(embedding_shape) == 2
        self.embedding_shape = embedding_shape
        n, c = embedding_shape
        self.register_buffer("y_embedding", torch.randn(n, c) / c ** 0.5)

    def forward(self, context: Dict):
        new_context = context.copy()
        emb = context["text_embeddings"]
        y = self.y_embedding.to(emb.device, emb.dtype)
        new_context["text_embeddings"] = y.unsqueeze(0).expand(emb.shape[0], -1, -1).contiguous()
        assert new_context["text_embeddings"].shape == emb.shape
        return new_context


class TextEmbeddingsAdapter(torch.nn.Module):
    """context["text_embeddings"] (B, L, C) -> (B, C, L) if ``swap_context_channels``, optionally projected
    (reference: context.py:115-140).  Timestep-invariant: evaluated once per sampling loop, not per step."""

    def __init__(self, swap_context_channels: bool = False, input_projection_dim: int = -1,
                 output_projection_dim: int = -1, **kwargs):
        super().__init__()
        self._swap_context_channels = swap_context_channels
        if output_projection_dim > 0 and input_projection_dim > 0:
            self._projection = torch.nn.Linear(input_projection_dim, output_projection_dim)
        else:
            self._projection = torch.nn.Identity()

    def forward(self, context: Dict):
        x = context["text_embeddings"]
        x = x.permute(0, 2, 1) if self._swap_context_channels else x
        return self._projection(x)
