"""Command-line mirror of DDPM/train.py for the two SalUn modes, on the sm_100a engine:

    python -m unlearn_saliency_b200.diffusion.cli --config cifar10_saliency_unlearn.yml --ckpt_folder results/cifar10/<ts> \
           --label_to_forget 0 --mode generate_mask
    python -m unlearn_saliency_b200.diffusion.cli --config cifar10_saliency_unlearn.yml --ckpt_folder results/cifar10/<ts> \
           --label_to_forget 0 --mode saliency_unlearn --mask_path results/cifar10/mask/0/with_0.5.pt --alpha 1e-3 --method rl

Same flags (train.py:14-93), same YAML keys, same directories: the config is read from ``configs/<name>`` (or the path
given), ``saliency_unlearn`` writes its snapshots and ``config.yaml`` under
``results/<dataset>/forget/<method>/<alpha>_<mask>/<timestamp>`` (functions/__init__.py:51-87), the mask goes to
``results/cifar10/mask/<label>/with_0.5.pt``.  ``--synthetic N`` replaces the CIFAR-10 loaders by N random images per
split (bring-up / benchmarks on a box without the dataset).  Other modes of train.py (train, forget, retrain) are the
reference's own training loops and are not served here.
"""
from __future__ import annotations

import argparse
import logging
import os
from types import SimpleNamespace

import numpy as np
import torch
import yaml


def dict2namespace(d):
    ns = SimpleNamespace()
    for k, v in d.items():
        setattr(ns, k, dict2namespace(v) if isinstance(v, dict) else v)
    return ns


def _load_config(path):
    if not os.path.exists(path):
        path = os.path.join("configs", path)
    with open(path) as f:
        return yaml.safe_load(f), path


def parse_args_and_config(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--config", type=str, required=True)
    p.add_argument("--ckpt_folder", type=str)
    p.add_argument("--mode", type=str, default="saliency_unlearn", help="generate_mask | saliency_unlearn")
    p.add_argument("--label_to_forget", type=int, default=0)
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--verbose", type=str, default="info")
    p.add_argument("--cond_scale", type=float, default=2.0)
    p.add_argument("--alpha", type=float, default=1.0)
    p.add_argument("--mask_path", type=str, default=None)
    p.add_argument("--method", type=str, default=None)
    p.add_argument("--mask_ratio", type=float, default=0.5,
                   help="accepted like the reference's flag, which generate_mask ignores too (threshold_list = [0.5], "
                        "runners/diffusion.py:1006)")
    # reference flags of the sampling / training modes: accepted so that the reference's command lines run unchanged
    p.add_argument("--sample_type", type=str, default="generalized")
    p.add_argument("--skip_type", type=str, default="uniform")
    p.add_argument("--timesteps", type=int, default=1000)
    p.add_argument("--eta", type=float, default=1.0)
    p.add_argument("--sequence", action="store_true")
    p.add_argument("--uc", type=bool, default=True)
    p.add_argument("--negative_guidance", type=float, default=7.5)
    p.add_argument("--sparse", type=bool, default=False)
    # additions of this mirror
    p.add_argument("--visualize", action="store_true",
                   help="sample_visualization at every snapshot of saliency_unlearn, like the reference (runners/diffusion.py:611-619)")
    p.add_argument("--precision", type=str, default=None, choices=["bf16", "split"],
                   help="engine build: default split (fp32-class) for generate_mask, bf16 for saliency_unlearn")
    p.add_argument("--synthetic", type=int, default=0, help="use N random images per split instead of CIFAR-10")
    args = p.parse_args(argv)
    cfg_dict, _ = _load_config(args.config)
    config = dict2namespace(cfg_dict)
    from datetime import datetime
    timestamp = datetime.now().strftime("%Y_%m_%d_%H%M%S")
    if args.mode == "saliency_unlearn":   # get_mask_config_and_setup_dirs, functions/__init__.py:51-87
        mp = args.mask_path or ""
        mask = next((k for k in ("origin", "inverted", "random", "without") if k in mp), "full")
        config.exp_root_dir = os.path.join("./results", config.data.dataset.lower(), "forget", str(args.method),
                                           f"{args.alpha}_{mask}", timestamp)
    else:                                 # get_config_and_setup_dirs / setup_dirs, functions/__init__.py:90-100
        config.exp_root_dir = os.path.join("./results", config.data.dataset.lower(), timestamp)
    config.log_dir = os.path.join(config.exp_root_dir, "logs")
    config.ckpt_dir = os.path.join(config.exp_root_dir, "ckpts")
    os.makedirs(config.log_dir, exist_ok=True)
    os.makedirs(config.ckpt_dir, exist_ok=True)
    with open(os.path.join(config.log_dir, "config.yaml"), "w") as f:
        yaml.safe_dump(cfg_dict, f, default_flow_style=False)
    level = getattr(logging, args.verbose.upper(), None)
    if not isinstance(level, int):
        raise ValueError(f"level {args.verbose} not supported")
    logging.basicConfig(level=level, format="%(levelname)s - %(filename)s - %(asctime)s - %(message)s",
                        handlers=[logging.StreamHandler(), logging.FileHandler(os.path.join(config.log_dir, "stdout.txt"))])
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(args.seed)
    return args, config


def _synthetic_loaders(n, config, label):
    from torch.utils.data import DataLoader, TensorDataset
    g = torch.Generator().manual_seed(0)
    S, bs = config.data.image_size, config.training.batch_size
    other = [k for k in range(config.data.n_classes) if k != label]
    cr = torch.tensor(other)[torch.randint(0, len(other), (n,), generator=g)]
    remain = TensorDataset(torch.rand(n, 3, S, S, generator=g), cr)
    forget = TensorDataset(torch.rand(n, 3, S, S, generator=g), torch.full((n,), label, dtype=torch.long))
    return DataLoader(remain, batch_size=bs, shuffle=True), DataLoader(forget, batch_size=bs, shuffle=True)


def main(argv=None):
    args, config = parse_args_and_config(argv)
    logging.info(f"Writing log file to {config.log_dir}")
    from .runner import Diffusion
    loaders = _synthetic_loaders(args.synthetic, config, args.label_to_forget) if args.synthetic else None
    runner = Diffusion(args, config, loaders=loaders)
    if args.mode == "generate_mask":                      # train.py:150-155
        runner.generate_mask()
    elif args.mode == "saliency_unlearn":
        runner.saliency_unlearn()
    elif args.mode == "visualization":
        runner.visualization()
    else:
        raise SystemExit(f"mode {args.mode!r}: generate_mask, saliency_unlearn and visualization run on the engine "
                         "(train / forget / retrain are the reference's own loops)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
