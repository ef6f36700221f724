"""Per-block error growth of the UNet / video UNet against the oracle (debugging aid, GPU)."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import nets as onets  # noqa: E402
from oracle import schedules  # noqa: E402
from tests.conftest import load_golden  # noqa: E402
from tests.helpers import fixture_state_dict, product_model, rel_l2  # noqa: E402


def main(name):
    fx = load_golden(name)
    m = product_model(fx)
    sd = fixture_state_dict(fx)
    p = fx["config"]["diffusion"]["score_network"]["params"]
    for i, s in fx["steps"].items():
        x = s["x"]
        B = x.shape[0]
        ot = {}
        if fx["kind"] == "unet3d":
            ref = onets.unet3d_forward(sd, p, x, s["logsnr_t"].expand(B), taps=ot)
            ctx = {"logsnr_t": s["logsnr_t"].expand(B).cuda().contiguous(), "timestep": torch.zeros(B).cuda()}
        else:
            t = torch.full((B,), i, dtype=torch.int64)
            ref = onets.unet_forward(sd, p, x, t, taps=ot)
            ctx = {"timestep": t.cuda()}
        pt = {}
        ctx["_taps"] = pt
        out = m.predict_score(x.cuda(), context=ctx)
        print(f"--- {name} step {i}: score rel_l2 = {rel_l2(out, ref):.4f}")
        for k, v in pt.items():
            r = ot[k]
            if r.dim() == 5:       # (B,C,F,H,W) -> (B*F,H,W,C)
                r = r.permute(0, 2, 3, 4, 1).reshape(-1, r.shape[3], r.shape[4], r.shape[1])
            else:
                r = r.permute(0, 2, 3, 1)
            print(f"   {k:10s} {rel_l2(v, r):.4f}")
        break


if __name__ == "__main__":
    for n in sys.argv[1:] or ["c1", "c5"]:
        main(n)
