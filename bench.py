"""Benchmark of the sampling hot path (BASELINE.json: images/sec over the full sampling loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload dit|unet]

A "step" is one full ``sample()`` call: the complete reverse-process loop (1000 timesteps) over one
batch of synthetic inputs.  Default workload = configs[1] of BASELINE.json: DiT on MNIST
(configs/image/mnist/dit.yaml), DDPM ancestral sampling with dynamic thresholding, batch 1024 per GPU
(weak scaling: global batch = 1024 * N), random-init weights with the zero-initialised tensors
re-randomised, Gaussian initial latents, in-kernel Philox step noise.  Prints ONE JSON line.

  value   whole-job images/s, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e     the same through the public API with HOST inputs: pinned x_T + labels copied H2D and the
          finished samples copied D2H inside the timed region, every step
  roofline  the dominant kernel (tcgen05 GEMM; the four contraction shapes of a DiT block), each
          launch timed alone with CUDA events after an L2 flush; algorithmic FLOPs / time vs the
          measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle (CPU port of the reference path) on the host cores, bounded sample
``--impl reference`` times that CPU port instead (the reference itself is Python and does not travel).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md section 8(d), 2*MAC per sample per network forward (pixart: x2 forwards under CFG; video: per clip)
FLOP_PER_IMAGE_STEP = {"dit": 0.7081e9, "unet": 10.684e9, "rf": 10.684e9, "pixart": 2 * 1.437e9, "video": 188.22e9}
FIXTURE = {"dit": "c2", "unet": "c1", "rf": "c3", "pixart": "c4", "video": "c5"}
DEFAULT_BATCH = {"dit": 1024, "unet": 64, "rf": 64, "pixart": 512, "video": 2}
SAMPLING_STEPS = 1000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["bf16_tflops_sustained"], d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def load_fixture(workload):
    from tests.conftest import load_golden
    return load_golden(FIXTURE[workload])


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every core the process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_port_images_per_sec(workload, batch, denoise_steps=2):
    """The oracle = CPU restatement of the reference path, all host threads, fp32.  Times
    `denoise_steps` reverse-process steps at the full batch and scales to the 1000-step loop (every
    step has identical cost)."""
    from tests.helpers import oracle_model
    use_all_host_threads()
    fx = load_fixture(workload)
    om = oracle_model(fx)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 1, 32, 32, generator=g)
    ctx = {"classes": torch.randint(0, 10, (batch,), generator=g)} if workload == "dit" else {}
    z = torch.randn(batch, 1, 32, 32, generator=g)

    def one(i, x):
        t = torch.full((batch,), i, dtype=torch.int64)
        o = om.score(x, t, ctx)
        from oracle import samplers as os_
        return os_.ancestral_discrete(x, o, z, i, om.tables, om.logvar, om.prediction, om.threshold)

    small = max(1, batch // 16)
    one(999, x[:small].clone()) if workload == "unet" else None       # touch the code path once (warm-up)
    t0 = time.perf_counter()
    for k in range(denoise_steps):
        x = one(SAMPLING_STEPS - 1 - k, x)
    dt = (time.perf_counter() - t0) / denoise_steps
    return batch / (dt * SAMPLING_STEPS), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()
    batch = args.batch
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_port_images_per_sec(args.workload, batch, 1)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt = cpu_port_images_per_sec(args.workload, batch, 1)
        vals.append(v)
    wall = time.perf_counter() - t_all
    value = sum(vals) / len(vals)
    sample = (f"1 of {SAMPLING_STEPS} reverse-process steps per bench step at batch {batch}, scaled x{SAMPLING_STEPS}")
    print(json.dumps({
        "impl": "reference", "metric": "images_per_sec_full_sampling_loop", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * batch / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall}))


def workload_config(args, n):
    name = {"dit": "DiT on MNIST (configs/image/mnist/dit.yaml), DDPM ancestral + dynamic thresholding",
            "unet": "DDPM UNet 32x32 (configs/image/mnist/ddpm_32x32_epsilon_discrete.yaml), ancestral",
            "rf": "Rectified flow UNet 32x32 (configs/image/mnist/rectified_flow_32x32.yaml), Euler",
            "pixart": "PixArt-alpha (configs/image/mnist/pixart_alpha.yaml), synthetic 77x768 text embeddings, CFG w=2",
            "video": "Video UNet-3D 16x32x32 (configs/video/moving_mnist/video_diffusion_models.yaml), v-pred"}[args.workload]
    return {"workload": name, "sampling_steps": args.sampling_steps or (1024 if args.workload == "video" else 1000),
            "per_gpu_batch": args.batch,
            "global_batch": args.batch * n, "parallelism": f"batch-sharded x{n}, one final all-gather",
            "l2": "activations per timestep exceed L2 (126 MB) at this batch; no flush between loops"}


def ncu_dram_bytes_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of one GEMM launch (qkv shape) from the committed
    `ncu --set full` summary under profiles/, or None."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_gemm_*_summary.csv")))
    if not files:
        return None
    try:
        rows = list(csv.reader(open(files[-1])))
        hdr = rows[0]
        rd = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum ["))
        wr = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum ["))
        scale = lambda h: {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[h.split("[")[1].rstrip("]")]
        return float(rows[1][rd]) * scale(hdr[rd]) + float(rows[1][wr]) * scale(hdr[wr])
    except Exception:
        return None


# ------------------------------------------------------------------------------------ roofline of the GEMM
def gemm_roofline(batch, device):
    """The four contractions of one DiT block at this batch WITH their real epilogues (bias; bias + GELU; bias + gate +
    fp32 residual in place).  `achieved` uses the in-loop condition: 20 launches per shape replayed from a CUDA graph
    (operands L2-resident as they are between the kernels of a timestep), CUDA events on the replay stream.  The
    cold number (one launch after an L2 flush, includes launch latency) is reported beside it."""
    from xdiffusion_b200 import ops
    M = batch * 16
    shapes = [(1152, 384, "qkv", False, 0), (384, 384, "proj", True, 0), (1536, 384, "fc1", False, ops.ACT_GELU),
              (384, 1536, "fc2", True, 0)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    tot_flop, tot_ms, per = 0.0, 0.0, {}
    for (n, k, name, rmw, act) in shapes:
        a = torch.randn(M, k, device=device).bfloat16()
        w = (torch.randn(n, k, device=device) * k ** -0.5).bfloat16()
        bias = torch.randn(n, device=device)
        if rmw:
            out = torch.randn(M, n, device=device)
            gate = torch.randn(M // 16, n, device=device) * 0.01
            call = lambda: ops.linear(a, w, bias, gate=gate, gate_rows=16, residual=out, out=out)
        else:
            out = torch.empty(M, n, device=device, dtype=torch.bfloat16)
            call = lambda: ops.linear(a, w, bias, act=act, out=out)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                call()
        graph.replay()
        torch.cuda.synchronize()
        reps = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            e1.synchronize()
            reps.append(e0.elapsed_time(e1) / 20)
        t = sorted(reps)[len(reps) // 2]
        cold = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            e1.synchronize()
            cold.append(e0.elapsed_time(e1))
        per[name] = {"us": round(t * 1e3, 2), "tflops": round(2.0 * M * n * k / t / 1e9, 1),
                     "us_cold_single_launch": round(sorted(cold)[2] * 1e3, 2)}
        tot_flop += 2.0 * M * n * k
        tot_ms += t
    return tot_flop / tot_ms / 1e9, per


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from tests.helpers import product_model
    from xdiffusion_b200 import ops
    from xdiffusion_b200.dist import gather_rows

    fx = load_fixture(args.workload)
    model = product_model(fx, device)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    shape = (B, 1, 16, 32, 32) if args.workload == "video" else (B, 1, 32, 32)
    x_host = torch.randn(shape, generator=g).pin_memory()
    cls_host = torch.randint(0, 10, (B,), generator=g).pin_memory()
    out_host = torch.empty(shape).pin_memory()
    x_dev, cls_dev = x_host.to(device), cls_host.to(device)
    extra, cfg_w = {}, None
    if args.workload == "pixart":
        from xdiffusion_b200.context import UnconditionalEmbeddingAdapter
        extra["text_embeddings"] = torch.randn(B, 77, 768, generator=g).to(device)
        model._unconditional_context = UnconditionalEmbeddingAdapter([77, 768]).to(device)
        cfg_w = 2.0
    n_steps = args.sampling_steps if args.sampling_steps else model.noise_scheduler().steps()

    def loop(x0, cls, seed):
        ctx = dict(extra)
        if args.workload in ("dit", "pixart"):
            ctx["classes"] = cls
        s, _ = model.sample(context=ctx, num_samples=B, initial_noise=x0, num_sampling_steps=n_steps, seed=seed,
                            classifier_free_guidance=cfg_w)
        return gather_rows(s, B * world) if world > 1 else s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l0 = ops.LAUNCHES
    loop(x_dev, cls_dev, 0)                       # builds + captures the step graph (untimed)
    launches_per_timestep = None
    if model._loops:
        lp = next(iter(model._loops.values()))
        ops.LAUNCHES = 0
        lp._step()                                 # one eager step = the kernels one graph replay launches
        launches_per_timestep = ops.LAUNCHES
    for w in range(max(args.warmup - 1, 0)):
        loop(x_dev, cls_dev, 1 + w)

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record()
            for k in range(args.steps):
                fn(k)
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, clk.summary()

    ms, clocks = timed(lambda k: loop(x_dev, cls_dev, 100 + k))

    def e2e_step(k):
        xd = x_host.to(device, non_blocking=True)
        cd = cls_host.to(device, non_blocking=True)
        out_host.copy_(loop(xd, cd, 200 + k)[:B], non_blocking=True)

    e2e_step(0)
    ms_e2e, _ = timed(e2e_step)
    total_images = B * world * args.steps
    value = total_images / (ms / 1e3)
    e2e_value = total_images / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    burst, sustained, hbm, src = measured_peaks()
    rl_tflops, per_shape = gemm_roofline(B, device) if args.workload == "dit" else (None, {})
    flop_img = FLOP_PER_IMAGE_STEP[args.workload] * n_steps
    line = {
        "metric": "images_per_sec_full_sampling_loop", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4 + cls_host.numel() * 8,
                "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": ((launches_per_timestep or 0) * n_steps + 2) * args.steps,
        "launches_per_timestep": launches_per_timestep,
        "ms_per_timestep": ms / args.steps / n_steps,
        "step_tensor_frac_of_sustained": value / world * flop_img / (sustained * 1e12),
    }
    if rl_tflops is not None:
        line["roofline"] = {"bound": "tensor", "achieved": rl_tflops, "peak": burst, "unit": "TFLOP/s",
                            "frac": rl_tflops / burst, "traffic": ncu_dram_bytes_per_launch(), "peak_source": src,
                            "kernel": "gemm_tc_kernel (tcgen05, CTA-pair 256x192 tiles, TMA epilogue): qkv+proj+fc1+fc2 of one DiT block with their real epilogues, 20 launches per shape replayed from a CUDA graph",
                            "per_shape": per_shape}
    if world == 1 and not args.no_cpu:
        v, dt = cpu_port_images_per_sec(args.workload, B, 2)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"2 of {SAMPLING_STEPS} reverse-process steps at batch {B} "
                                          f"({dt:.2f} s/step), scaled x{SAMPLING_STEPS}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dit", choices=list(FIXTURE))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 1024 DiT, 64 UNet)")
    ap.add_argument("--sampling-steps", type=int, default=0, help="0 = the scheduler's full step count")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = DEFAULT_BATCH[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
