"""Stable Diffusion side of SalUn: the loop bodies of SD/train-scripts/train-esd.py,
random_label.py and generate_mask.py around a LatentDiffusion-like model, with the per-step mask multiply (a 6.9 GB int64
H2D copy per step in the reference), Adam, the saliency accumulation (a 3.4 GB D2H copy per batch) and the top-k on the
sm_100a tail kernels.  The 860 M-parameter U-Net forward / backward (row a16) runs through PyTorch."""
from .loops import SDTail, certain_label, esd_iteration, generate_mask, select_parameters, train_esd  # noqa: F401
