"""The product package's pure-Python parameter table (diffusion/config.py: unet_param_table) against the PyTorch
restatement of Conditional_Model (oracle/unet.py, itself pinned to the reference's 334 named_parameters() keys by
tests/golden/ddpm_tiny.npz) -- names, order and shapes."""
import numpy as np
import os

from oracle.unet import ConditionalUNet
from tests.golden.make_golden_ddpm import small_config, tiny_config
from unlearn_saliency_b200.diffusion.config import cifar10_config, unet_param_table


def test_param_table_matches_torch_module_and_reference_keys():
    for cfg in (tiny_config(), small_config(), cifar10_config()):
        want = [(k, tuple(p.shape)) for k, p in ConditionalUNet(cfg).named_parameters()]
        assert list(unet_param_table(cfg).items()) == want
    t = unet_param_table(cifar10_config())
    assert len(t) == 334 and sum(int(np.prod(s)) for s in t.values()) == 38632323
    strip = lambda ks: [str(k)[7:] if str(k).startswith("module.") else str(k) for k in ks]
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm_tiny.npz"))
    assert strip(z["keys_full"]) == list(t.keys())                       # the reference's own key list (make_golden_ddpm.py)
    assert strip(z["keys_tiny"]) == list(unet_param_table(tiny_config()).keys())
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm_small.npz"))
    assert strip(z["keys"]) == list(unet_param_table(small_config()).keys())
