"""Training-time timestep samplers are out of scope; the discrete scheduler's YAML still names one
(reference: scheduler.py:158), so an inert stand-in keeps the constructor signature working."""


class UniformSampler:
    def __init__(self, num_timesteps: int = 0, **kwargs):
        self.num_timesteps = num_timesteps

    def sample(self, *a, **k):
        raise NotImplementedError("timestep importance sampling is a training feature (out of scope)")
