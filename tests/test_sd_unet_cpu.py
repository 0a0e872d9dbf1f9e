"""The host mirror's parameter table of the SD U-Net (sd/engine.py) against the reference's named_parameters() keys
(tests/golden/sd_unet.npz, generated from the unmodified UNetModel by make_golden_sd.py)."""
import math
import os

import numpy as np

from tests.golden.make_golden_sd import CONFIGS
from unlearn_saliency_b200.sd.engine import sd_unet_param_table, sd_v1_config

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "sd_unet.npz"))


def test_param_tables_match_reference_keys():
    for tag, c in CONFIGS.items():
        assert [str(k) for k in Z[f"{tag}_keys"]] == list(sd_unet_param_table(c["cfg"]).keys())
    t = sd_unet_param_table(sd_v1_config())
    assert [str(k) for k in Z["v14_keys"]] == list(t.keys())
    assert len(t) == 686 and sum(math.prod(s) for s in t.values()) == int(Z["v14_numel"]) == 859520964


def test_config_from_state_dict_round_trips():
    import torch
    from unlearn_saliency_b200.sd.engine import config_from_state_dict, sd_unet_param_table, sd_v1_config
    from tests.golden.make_golden_sd import CONFIGS
    for cfg in [sd_v1_config()] + [c["cfg"] for c in CONFIGS.values()]:
        table = sd_unet_param_table(cfg)
        fake = {"model.diffusion_model." + k: torch.empty(shp, device="meta") for k, shp in table.items()}
        got = config_from_state_dict(fake, num_heads=cfg["num_heads"])
        want = {**cfg, "transformer_depth": cfg.get("transformer_depth", 1), "attention_resolutions": sorted(cfg["attention_resolutions"])}
        assert got == want, (got, want)       # the order of attention_resolutions carries no meaning (membership test)
    fake.pop("model.diffusion_model.out.2.bias")
    import pytest
    with pytest.raises(ValueError):
        config_from_state_dict(fake)
