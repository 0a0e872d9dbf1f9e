"""Host logic of the SD DDIM sampler mirror (no GPU): schedule tables equal the restatement of
SD/ldm/modules/diffusionmodules/util.py:20-96, and the package exports what INTEGRATION.md names."""
import numpy as np
import pytest

from oracle import sd_unet as OS


def test_schedule_tables_equal_restatement():
    from unlearn_saliency_b200.sd import sampler as SP
    ac = np.cumprod(1.0 - SP.make_beta_schedule())
    np.testing.assert_array_equal(ac, OS.sd_alphas_cumprod())
    for n, eta in ((50, 0.0), (10, 0.7), (25, 1.0), (200, 0.3)):
        steps = SP.make_ddim_timesteps("uniform", n, 1000)
        sg, a, ap = SP.make_ddim_sampling_parameters(ac, steps, eta)
        rs, ra, rap, rsg = OS.ddim_schedule(ac, n, eta)
        for mine, ref in ((steps, rs), (a, ra), (ap, rap), (sg, rsg)):
            np.testing.assert_array_equal(mine, ref)
    assert SP.make_ddim_timesteps("quad", 10, 1000)[0] == 1
    with pytest.raises(NotImplementedError):
        SP.make_ddim_timesteps("cosine", 10, 1000)
    with pytest.raises(ValueError):
        SP.make_beta_schedule("cosine")


def test_sd_package_exports():
    import unlearn_saliency_b200.sd as sd
    for name in ("train_esd", "certain_label", "generate_mask", "esd_iteration", "select_parameters", "SDTail", "SDUNetEngine",
                 "EngineDDIMSampler", "EngineApplyModel", "make_quick_sample_till_t", "sample_model", "sd_v1_config",
                 "sd_unet_param_table"):
        assert hasattr(sd, name), name
