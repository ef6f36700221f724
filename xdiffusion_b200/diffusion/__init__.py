from .ddpm import DiffusionModel, GaussianDiffusion_DDPM, PredictionType  # noqa: F401
