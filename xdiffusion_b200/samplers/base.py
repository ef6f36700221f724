"""Sampler interface kept from the reference (samplers/base.py:6-33) plus the shared fused-step driver."""
from typing import Dict, Optional

import torch

from .. import ops


class ReverseProcessSampler:
    #: "philox": Gaussian noise is generated inside the fused step kernel (counter = element, step; key = seed)
    #: "torch":  torch.randn_like(x) per step, i.e. the reference's RNG consumption (samplers/ancestral.py:59)
    noise_source = "philox"
    seed = 0

    def p_sample(self, x: torch.Tensor, context: Dict, unconditional_context: Optional[Dict], diffusion_model,
                 guidance_fn=None, classifier_free_guidance: Optional[float] = None):
        raise NotImplementedError

    # ------------------------------------------------------------------ helpers shared by the samplers
    @staticmethod
    def _score(diffusion_model, x, context, unconditional_context, classifier_free_guidance):
        """eps/v prediction, with the classifier-free-guidance combine of ancestral.py:211-238.
        (The reference's own CFG call omits `diffusion_model` and raises TypeError, SURVEY.md section 4;
        the intended computation is implemented.)"""
        if diffusion_model.is_learned_sigma():
            raise NotImplementedError("learned-sigma score networks are outside the covered hot path")
        cfg = classifier_free_guidance if classifier_free_guidance is not None \
            else diffusion_model.classifier_free_guidance()
        both = context.get("_cfg_both")
        if cfg >= 0.0 and unconditional_context is not None and both is not None:
            # the sampling loop prepared [conditional | unconditional] conditioning for 2B rows (diffusion/ddpm.py,
            # _DeviceLoop._merge; identity input preprocessing only): one forward, rows B.. are the unconditional scores
            B = x.shape[0]
            o2 = diffusion_model.predict_score(torch.cat([x, x], 0), context=both)
            out = torch.empty_like(x)
            torch.ops.xdb200.cfg_combine(o2[:B], o2[B:], float(cfg), out)
            return out
        xin = diffusion_model.process_input(x=x, context=context)
        o = diffusion_model.predict_score(xin, context=context)
        if cfg >= 0.0 and unconditional_context is not None:
            ou = diffusion_model.predict_score(
                diffusion_model.process_input(x=x, context=unconditional_context), context=unconditional_context)
            out = torch.empty_like(o)
            torch.ops.xdb200.cfg_combine(o.contiguous(), ou.contiguous(), float(cfg), out)
            o = out
        return o

    def _launch_step(self, mode, x, o, context, diffusion_model, sampler_name):
        sched = diffusion_model.noise_scheduler()
        N = context.get("num_sampling_steps", sched.steps())
        coefs, form = diffusion_model.step_coefficients(sampler_name, N)
        idx = context["timestep_idx"]
        idx_dev, idx_host = (idx, -1) if torch.is_tensor(idx) else (None, int(idx))
        thr = diffusion_model.dynamic_threshold(x[0].numel())
        z, z_stride = context.get("noise"), 0
        if z is not None and z.dim() == x.dim() + 1:          # [N, B, ...] per-step noise table
            z_stride = x.numel()
        elif z is None and mode == ops.MODE_ANCESTRAL and self.noise_source == "torch":
            z = torch.randn_like(x)
        out = context.get("out")
        if out is None:
            out = torch.empty_like(x)
        pred_v = int(diffusion_model.prediction_type().name == "V")
        torch.ops.xdb200.sampler_step(mode, form, pred_v, x, o, z, z_stride, out, coefs, idx_dev, idx_host,
                                      int(thr is not None), thr[0] if thr else 0, thr[1] if thr else 0.0,
                                      thr[2] if thr else 0.0, int(context.get("seed", self.seed)),
                                      context.get("seed_dev"), int(context.get("row_offset", 0)) * x[0].numel())
        return out
