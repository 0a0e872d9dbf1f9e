"""mirror of Classification/unlearn/__init__.py:22-61 for the methods on the SalUn hot path."""
from .FT import FT, FT_l1
from .GA import GA
from .RL import RL
from .impl import iterative_unlearn  # noqa: F401


def raw(data_loaders, model, criterion, args, mask=None):
    pass


def get_unlearn_method(name):
    """method usage: function(data_loaders, model, criterion, args[, mask])"""
    if name == "raw":
        return raw
    if name == "RL":
        return RL
    if name == "GA":
        return GA
    if name == "FT":
        return FT
    if name == "FT_l1":
        return FT_l1
    raise NotImplementedError(f"Unlearn method {name} is not on the sm_100a hot path (served: raw, RL, GA, FT, FT_l1)")
