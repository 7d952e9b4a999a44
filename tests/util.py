import torch


def rel_err(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    den = b.norm().item()
    if den == 0:
        return (a - b).norm().item()
    return (a - b).norm().item() / den


def cfg_from_flags(F):
    """Namespace for the oracle carrying the current FLAGS values."""
    from oracle.config import default_cfg, _DEFAULTS
    return default_cfg(**{k: getattr(F, k) for k in _DEFAULTS})
