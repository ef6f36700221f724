"""LoRA fixture: the REAL reference with its own LoRA injection (xdiffusion/lora.py:228-343).

    python tests/golden/make_lora.py            # writes tests/golden/lora_c1.pt and lora_c7.pt

Flow, mirroring sampling/image/mnist/sample.py:86-98: the reference model gets the fixture's synthetic weights; a LoRA
file (flat list [up_0, down_0, ...] of Parameters, rank 4, seeded NON-zero values -- the reference initialises ``up`` to
zero, which would make the comparison vacuous) is written in the order of the reference's ``save_lora_weights``, loaded
back through the reference's ``load_lora_weights`` and the score is evaluated with the injected LoraInjected* modules.
The fixture stores the (up, down) shapes in file order, the seed, the inputs and the reference scores (with and without
LoRA).  Only runs in the authoring container.
"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.golden.make_golden import build  # noqa: E402
from xdiffusion_b200.lora import synth_lora  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RANK, SEED = 4, 5


def make(name):
    model, kind, manifest = build(name)                   # (bootstraps the reference import path)
    from xdiffusion.lora import extract_lora_ups_down, inject_trainable_lora, load_lora_weights
    size = model.config().to_dict()["diffusion"]["sampling"]["output_spatial_size"]
    g = torch.Generator().manual_seed(99)
    B = 2
    x = torch.randn(B, 1, size, size, generator=g)
    ctx = {"timestep": torch.tensor([500, 37])}
    if name == "c7":
        ctx["text_embeddings"] = torch.randn(B, 77, 768, generator=g)
        ctx["text_prompts"] = ["", ""]
    with torch.no_grad():
        base = model.predict_score(x, context=dict(ctx))
    # the file order and shapes come from the reference's own traversal: inject into a scratch copy, read the shapes
    scratch, _, _ = build(name)
    inject_trainable_lora(scratch, r=RANK)
    shapes = [(tuple(u.weight.shape), tuple(d.weight.shape)) for u, d in extract_lora_ups_down(scratch)]
    loras = synth_lora(shapes, seed=SEED)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "lora.pt")
        torch.save(loras, path)
        load_lora_weights(model, lora_path=path)          # the reference's loader, on the real model
    model.eval()
    with torch.no_grad():
        with_lora = model.predict_score(x, context=dict(ctx))
    out = {"name": name, "rank": RANK, "seed": SEED, "shapes": shapes, "x": x,
           "ctx": {k: v for k, v in ctx.items() if torch.is_tensor(v)}, "score_base": base, "score_lora": with_lora}
    torch.save(out, os.path.join(HERE, f"lora_{name}.pt"))
    print(name, len(shapes), "LoRA layers; |delta score| / |score| =",
          float((with_lora - base).norm() / base.norm()))


if __name__ == "__main__":
    for n in sys.argv[1:] or ["c1", "c7"]:
        make(n)
