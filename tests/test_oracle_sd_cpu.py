"""oracle/sd_unet.py (the torch restatement of the SD U-Net the GPU tests and the eager baseline use) against eps of the
UNMODIFIED reference UNetModel (tests/golden/sd_unet.npz)."""
import os

import numpy as np
import torch

from oracle import sd_unet as OS
from tests.golden.make_golden_sd import CONFIGS, sd_inputs, sd_synth_weights
from unlearn_saliency_b200.sd.engine import sd_unet_param_table

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "sd_unet.npz"))


def test_unet_forward_equals_reference():
    torch.set_num_threads(8)
    for tag, c in CONFIGS.items():
        P = sd_synth_weights(sd_unet_param_table(c["cfg"]), seed=7)
        x, t, ctx = sd_inputs(tag)
        with torch.no_grad():
            eps = OS.unet_forward(P, c["cfg"], x, t, ctx)
        np.testing.assert_allclose(eps.numpy(), Z[f"{tag}_eps"], rtol=2e-4, atol=2e-5)


def test_ddim_schedule_values():
    ac = OS.sd_alphas_cumprod()
    steps, a, ap, sg = OS.ddim_schedule(ac, 50, 0.0)
    assert steps[0] == 1 and steps[-1] == 981 and len(steps) == 50 and np.all(sg == 0)
    assert abs(ac[0] - (1 - 0.00085)) < 1e-9 and abs(ap[0] - ac[0]) < 1e-12 and abs(a[1] - ac[21]) < 1e-12
