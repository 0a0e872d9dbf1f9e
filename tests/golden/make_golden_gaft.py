"""Golden fixtures for the GA / FT / FT_l1 entry points, from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_gaft.py  ->  tests/golden/resnet18_gaft.npz

unlearn.GA (GA.py:107-150), unlearn.FT and unlearn.FT_l1 (FT.py:116-180) are run for one epoch (2 batches of 16) and
for a single batch on resnet18 with the 0.5 mask of resnet18_mask.npz; sampled final parameters, per-tensor norms and a
BatchNorm buffer are stored.  Same inputs as make_golden.py (seed 11 / 13)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, ref_args, sample_idx  # noqa: E402


def main():
    torch.set_num_threads(8)
    gm, unlearn, model_dict = import_reference()
    from oracle import classification as OC
    crit = torch.nn.CrossEntropyLoss()
    params, buffers = OC.synth_state(10, seed=0)
    sd = OC.state_dict_of(params, buffers)
    n = 11173962
    zm = np.load(os.path.join(HERE, "resnet18_mask.npz"))
    flat_mask = torch.from_numpy(np.unpackbits(zm["bits_0.5"], bitorder="little")[:n].astype(np.int64))
    mask = OC.split_mask(flat_mask, OC.resnet18_param_shapes(10))
    g = torch.Generator().manual_seed(11)
    x = torch.rand(32, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (32,), generator=g)
    g = torch.Generator().manual_seed(13)
    xr = torch.rand(32, 3, 32, 32, generator=g)
    yr = torch.randint(0, 10, (32,), generator=g)
    mk = lambda a, b, k: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(a[:k], b[:k]), batch_size=16, shuffle=False)
    res = {}
    for name, kw in (("GA", {}), ("FT", {}), ("FT_l1", dict(alpha=5e-4, unlearn_epochs=2, no_l1_epochs=0))):
        for tag, k in (("1", 16), ("2", 32)):
            model = model_dict["resnet18"](num_classes=10)
            model.load_state_dict(sd)
            loaders = {"forget": mk(x, y, k), "retain": mk(xr, yr, k)}
            a = ref_args("/tmp", unlearn=name, **kw)
            acc = getattr(unlearn, name)(loaders, model, crit, a, mask)
            fin = dict(model.named_parameters())
            res[f"{name}_{tag}_psample"] = np.concatenate(
                [p.detach().flatten()[sample_idx(p.numel())].numpy() for p in fin.values()])
            res[f"{name}_{tag}_pnorm"] = np.array([p.detach().norm().item() for p in fin.values()], dtype=np.float64)
            res[f"{name}_{tag}_rm_bn1"] = dict(model.named_buffers())["bn1.running_mean"].numpy().copy()
            res[f"{name}_{tag}_acc"] = np.float64(acc)
    np.savez_compressed(os.path.join(HERE, "resnet18_gaft.npz"), **res)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in res.items() if k.endswith("acc")})


if __name__ == "__main__":
    main()
