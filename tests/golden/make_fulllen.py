"""Full-length golden runs: the REAL reference (imported from /root/reference) sampling the complete reverse
process of the BASELINE configurations on a few rows, with seeded x_T and per-step noise.

    python tests/golden/make_fulllen.py [c1 c2 c3]      # writes tests/golden/full_<name>.pt

Rows are independent on this path (per-sample norms, attention, quantile), so a GPU run at the benchmark batch
size whose first rows carry the same x_T / noise / conditioning must reproduce these samples
(tests/test_fulllen_gpu.py).  Nothing but the final samples is stored: x_T and the noise are regenerated from the
seed by ``fulllen_inputs`` (torch's CPU generator), a probe of both is kept to detect RNG drift.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

HERE = os.path.dirname(os.path.abspath(__file__))

#: name -> (rows, seed); the step count is the config's own (1000)
RUNS = {"c1": (4, 101), "c2": (8, 102), "c3": (4, 103)}


def fulllen_inputs(name, rows, seed, steps):
    """x_T [rows,1,32,32], noise [steps, rows,1,32,32] (row = loop index), classes or None."""
    g = torch.Generator().manual_seed(seed)
    x_T = torch.randn(rows, 1, 32, 32, generator=g)
    noise = torch.randn(steps, rows, 1, 32, 32, generator=g)
    classes = torch.randint(0, 10, (rows,), generator=g) if name == "c2" else None
    return x_T, noise, classes


def main(names):
    from make_golden import NoiseFeeder, build
    from tests.helpers import oracle_model
    from tests.conftest import load_golden
    for name in names:
        rows, seed = RUNS[name]
        model, kind, _ = build(name)
        N = model.noise_scheduler().steps()
        x_T, noise, classes = fulllen_inputs(name, rows, seed, N)
        ctx = {} if classes is None else {"classes": classes}
        t0 = time.time()
        with NoiseFeeder() as feeder:
            feeder.queue = [noise[i].clone() for i in reversed(range(N))]
            samples, _ = model.sample(context=dict(ctx), num_samples=rows, num_sampling_steps=N,
                                      initial_noise=x_T.clone())
        t_ref = time.time() - t0
        om = oracle_model(load_golden(name))
        ora = om.sample(x_T, [noise[i] for i in range(N)], ctx=dict(ctx), num_sampling_steps=N)
        d = float((ora - samples).norm() / samples.norm())
        print(f"{name}: reference {t_ref:.0f} s, oracle vs reference rel L2 after {N} steps = {d:.3e}", flush=True)
        torch.save({"name": name, "rows": rows, "seed": seed, "steps": N, "samples": samples,
                    "oracle_vs_reference": d,
                    "probe": {"x_T": x_T.flatten()[:8].clone(), "noise": noise.flatten()[-8:].clone()}},
                   os.path.join(HERE, f"full_{name}.pt"))


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    main(sys.argv[1:] or list(RUNS))
