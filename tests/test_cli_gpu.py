"""Drop-in entry points end to end on synthetic data: generate_mask -> main_random (RL) with the on-disk formats of
the reference (generate_mask.py:82 mask files; impl.py:21-30 checkpoint)."""
import os

import pytest
import torch

from oracle import classification as OC

pytestmark = pytest.mark.gpu


def test_generate_mask_then_rl(tmp_path):
    from unlearn_saliency_b200.classification import cli
    params, buffers = OC.synth_state(10, seed=0)
    ckpt = tmp_path / "model.pth.tar"
    torch.save({"state_dict": OC.state_dict_of(params, buffers)}, ckpt)
    mdir, odir = tmp_path / "mask", tmp_path / "out"
    common = ["--synthetic", "640", "--num_indexes_to_replace", "128", "--batch_size", "64", "--model_path", str(ckpt),
              "--seed", "2"]
    cli.main(["generate_mask", "--save_dir", str(mdir), "--unlearn_epochs", "1"] + common)
    n = 11173962
    shapes = OC.resnet18_param_shapes(10)
    for r in OC.THRESHOLDS:
        m = torch.load(mdir / f"with_{r}.pt")
        assert list(m.keys()) == list(shapes.keys())                       # named_parameters order, bare names
        assert all(v.dtype == torch.int64 and v.is_cuda and tuple(v.shape) == shapes[k] for k, v in m.items())
        assert sum(int(v.sum()) for v in m.values()) == int(n * r)          # k = int(N * ratio), generate_mask.py:60
    m01, m05 = torch.load(mdir / "with_0.1.pt"), torch.load(mdir / "with_0.5.pt")
    assert all(bool((m05[k] >= m01[k]).all()) for k in m01)                  # nested top-k sets
    cli.main(["main_random", "--unlearn", "RL", "--unlearn_epochs", "1", "--unlearn_lr", "0.013", "--save_dir", str(odir),
              "--mask_path", str(mdir / "with_0.5.pt"), "--print_freq", "2"] + common)
    ck = torch.load(odir / "RLcheckpoint.pth.tar")
    sd = ck["state_dict"]
    assert set(OC.state_dict_of(params, buffers).keys()) == set(sd.keys())
    assert int(sd["bn1.num_batches_tracked"]) == 10                          # 2 forget + 8 retain steps
    moved = kept = 0
    for k, p in params.items():
        new, msk = sd[k].cpu(), m05[k].cpu()
        assert torch.equal(new[msk == 0], p[msk == 0]), k                    # masked-out weights restored exactly
        moved += int((new[msk == 1] != p[msk == 1]).sum())
        kept += int((msk == 0).sum())
    assert moved > 0.9 * (n - kept)
    assert "accuracy" in ck["evaluation_result"] and set(ck["evaluation_result"]["accuracy"]) == {"retain", "forget", "val", "test"}
    # methods outside the hot path fail loudly
    with pytest.raises(NotImplementedError):
        cli.main(["main_forget", "--unlearn", "wfisher", "--save_dir", str(odir)] + common)
