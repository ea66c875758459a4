"""A prior that may only be filled from external samples (reference config/placeholder_sample.py:6-33)."""

from ..infer.distributions import Distribution


class SamplePlaceholderError(Exception):
    """Raised when a PlaceholderSample is sampled outside a substitute / Predictive context."""


class PlaceholderSample(Distribution):
    raises_on_sample = True

    def __init__(self):
        pass

    def _params(self):
        import torch
        return (torch.zeros((), dtype=torch.float64),)

    def sample(self, key=None, sample_shape=()):
        raise SamplePlaceholderError(
            "Attempted to sample a PosteriorSample parameter outside of a Predictive() context. This likely "
            "means you did not provide posterior samples to the context via Predictive() or substitute().")
