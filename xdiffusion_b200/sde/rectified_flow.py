"""RectifiedFlow(N, T): the three scalars the Euler sampler needs (reference: sde/rectified_flow.py:4-28)."""


class RectifiedFlow:
    def __init__(self, N, T, **kwargs):
        self._N, self._T = N, T

    @property
    def T(self):
        return self._T

    @property
    def N(self):
        return self._N

    def sigma_t(self, t: float):
        return 0.0

    def noise_scale(self) -> float:
        return 1.0
