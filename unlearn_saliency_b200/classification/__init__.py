"""Host-side mirror of the reference's Classification/ interface for the SalUn hot path.

Same function names, argument meaning and on-disk formats as
  Classification/generate_mask.py   (save_gradient_ratio)
  Classification/unlearn/           (get_unlearn_method, RL, GA, FT, raw; iterative_unlearn)
with ``model`` being (or being converted to) a :class:`unlearn_saliency_b200.engine.ResNetEngine`.
"""
from .common import as_engine, check_criterion  # noqa: F401
