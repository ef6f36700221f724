"""The sampling loop (oracle).  TEST INFRASTRUCTURE ONLY.

Restates GaussianDiffusion_DDPM.sample()/_p_sample_loop()
(xdiffusion/diffusion/ddpm.py:544-669, 866-987) for the five benchmark configs,
with the initial latent and the per-step noise injected so that runs are
comparable across devices.
"""
import torch

from . import nets, samplers, schedules


class OracleModel:
    """kind in {"unet", "dit", "pixart", "unet3d"}; cfg = full YAML dict."""

    def __init__(self, kind, cfg, sd):
        self.kind, self.cfg, self.sd = kind, cfg, sd
        d = cfg["diffusion"]
        self.p = d["score_network"]["params"]
        self.prediction = d["parameterization"]
        ns = d["noise_scheduler"]
        self.sched_kind = ns["target"].rsplit(".", 1)[1]
        sp = ns["params"]
        if self.sched_kind == "DiscreteNoiseScheduler":
            self.tables = schedules.discrete_tables(sp["num_scales"], sp["schedule_type"],
                                                    sp.get("min_beta", 1e-4), sp.get("max_beta", 2e-2))
            self.logvar = schedules.fixed_large_logvar(self.tables)
            self.steps = sp["num_scales"]
        elif self.sched_kind == "ContinuousNoiseScheduler":
            fn = schedules.cosine_logsnr_table if sp["logsnr_schedule"] == "cosine" else schedules.linear_logsnr_table
            self.gammas = fn(sp["num_scales"], sp["logsnr_min"], sp["logsnr_max"])
            self.steps = sp["num_scales"]
        else:
            self.steps = sp["steps"]
            self.N = d["sde"]["params"]["N"]
            self.T = d["sde"]["params"]["T"]
        sr = cfg.get("super_resolution")                  # cascade stage: low-resolution conditioning (ddpm.py:613-618)
        self.sr = None if sr is None else {"key": sr["conditioning_key"], "size": sr["super_resolution_size"],
                                           "level": sr.get("sampling_augmentation_level")}
        dt = d.get("dynamic_thresholding")
        self.threshold = (dt["p"], dt["c"]) if dt and dt.get("enable") else None

    @torch.no_grad()
    def score(self, x, t, ctx):
        if self.kind == "unet":
            tin = ctx["logsnr_t"] if nets.time_input_key(self.p) == "logsnr_t" else t
            return nets.unet_forward(self.sd, self.p, x, tin, text=ctx.get("text_embeddings"))
        if self.kind == "effunet":
            # InputPreprocessor + GaussianConditioningAugmentationToTimestep (layers/super_resolution.py): the augmentation
            # timestep is (ones(B, long) * steps * level).to(long); ctx["z_cond"] is the noise q_sample draws at this call
            B = x.shape[0]
            s = (torch.ones(B, dtype=torch.long) * self.steps * self.sr["level"]).to(torch.long)
            xin = nets.sr_input(x, ctx[self.sr["key"]], self.sr["size"], self.tables, s, ctx["z_cond"])
            return nets.effunet_forward(self.sd, self.p, xin, t, text=ctx.get("text_embeddings"), augmentation_timestep=s)
        if self.kind == "dit":
            return nets.dit_forward(self.sd, self.p, x, t, ctx["classes"])
        if self.kind == "pixart":
            return nets.pixart_forward(self.sd, self.p, x, t, ctx["text_embeddings"])
        if self.kind == "unet3d":
            return nets.unet3d_forward(self.sd, self.p, x, ctx["logsnr_t"])
        raise NotImplementedError(self.kind)

    @torch.no_grad()
    def sample(self, x_T, noises, ctx=None, num_sampling_steps=None, sampler="ancestral",
               cfg_scale=None, uncond_ctx=None, trace=None, cond_noises=None):
        """noises[i] is the z used at loop index i (ignored by ddim / euler).  Returns the
        un-normalised samples in [0,1] (ddpm.py:667).  ``trace`` (list) receives
        (i, score, x_next) per step."""
        ctx = dict(ctx or {})
        N = num_sampling_steps or self.steps
        B = x_T.shape[0]
        x = x_T
        mask = ctx.get("video_mask")                      # (B, F) bool: True = generate, False = keep ctx["x0"]
        if mask is not None:
            mask = mask[:, None, :, None, None]
        for i in reversed(range(N)):
            if mask is not None:                          # ddpm.py:963-967
                x = torch.where(mask, x, ctx["x0"])
            c = dict(ctx)
            if cond_noises is not None:                   # super-resolution stage: conditioning noise of this step
                c["z_cond"] = cond_noises[i]
            if self.sched_kind == "ContinuousNoiseScheduler":
                idx_s, idx_t = schedules.continuous_indices(i, N, self.steps)
                lam_s, lam_t = self.gammas[idx_s], self.gammas[idx_t]
                c["logsnr_t"] = lam_t.expand(B)
                t = torch.full((B,), schedules.continuous_time(i, N), dtype=torch.float32)
            elif self.sched_kind == "DiscreteNoiseScheduler":
                t = torch.full((B,), i, dtype=torch.int64)
            else:
                t = torch.full((B,), schedules.rectified_flow_time(i, self.N, self.T), dtype=torch.float32)
            o = self.score(x, t, c)
            if cfg_scale is not None and uncond_ctx is not None:
                cu = dict(uncond_ctx)
                if "logsnr_t" in c:
                    cu["logsnr_t"] = c["logsnr_t"]
                o = samplers.cfg_combine(o, self.score(x, t, cu), cfg_scale)
            if self.sched_kind == "DiscreteRectifiedFlowNoiseScheduler":
                x_next = samplers.euler_flow(x, o, self.N)
            elif self.sched_kind == "DiscreteNoiseScheduler":
                x_next = samplers.ancestral_discrete(x, o, noises[i], i, self.tables, self.logvar,
                                                     self.prediction, self.threshold)
            elif sampler == "ddim":
                x_next = samplers.ddim_continuous(x, o, i, lam_s, lam_t, self.prediction, self.threshold)
            else:
                x_next = samplers.ancestral_continuous(x, o, noises[i], i, lam_s, lam_t,
                                                       self.prediction, self.threshold)
            if trace is not None:
                trace.append((i, o, x_next))
            x = x_next
            if mask is not None:                          # ddpm.py:979-982
                x = torch.where(mask, x, ctx["x0"])
        return samplers.unnormalize(x)
