"""Inert stand-in (random-forest propensity models are out of scope)."""


def reset_random_seeds(seed):
    return None


Ensemble = object
ECV = object
