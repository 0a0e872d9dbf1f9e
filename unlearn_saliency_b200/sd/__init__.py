"""Stable Diffusion side of SalUn: the loop bodies of SD/train-scripts/train-esd.py,
random_label.py and generate_mask.py around a LatentDiffusion-like model, with the per-step mask multiply (a 6.9 GB int64
H2D copy per step in the reference), Adam, the saliency accumulation (a 3.4 GB D2H copy per batch) and the top-k on the
sm_100a tail kernels.  The U-Net forward (row a16) has its own engine (sd/engine.py: the op-level C ABI of csrc/salun_ops.cu replayed from a CUDA
graph) and the DDIM partial sampler of ESD runs on it (sd/sampler.py, row f1); the one pass per iteration that needs autograd
(e_n of the trained model) stays the caller's module."""
from .loops import SDTail, certain_label, esd_iteration, generate_mask, select_parameters, train_esd  # noqa: F401
from .engine import SDUNetEngine, sd_unet_param_table, sd_v1_config  # noqa: F401
from .sampler import EngineApplyModel, EngineDDIMSampler, make_quick_sample_till_t, sample_model  # noqa: F401
