"""Entry points with the reference's command lines:

    python -m unlearn_saliency_b200.classification.cli generate_mask --save_dir M --model_path CKPT \
           --num_indexes_to_replace 4500 --unlearn_epochs 1                        (Classification/generate_mask.py:85-202)
    python -m unlearn_saliency_b200.classification.cli main_random  --unlearn RL --unlearn_epochs 10 --unlearn_lr 0.013 \
           --num_indexes_to_replace 4500 --model_path CKPT --save_dir OUT --mask_path M/with_0.5.pt   (main_random.py:15-191)
    python -m unlearn_saliency_b200.classification.cli main_forget  --unlearn FT ...                    (main_forget.py)
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import torch

from ..engine import ResNetEngine
from . import arg_parser
from .data import make_loaders
from .evaluation import SVC_MIA, validate
from .generate_mask import save_gradient_ratio
from .unlearn import get_unlearn_method


def _engine_from_args(args) -> ResNetEngine:
    eng = ResNetEngine(args.arch, args.num_classes, args.input_size, max_batch=args.batch_size,
                       device=f"cuda:{int(args.gpu)}")
    if args.model_path:
        ckpt = torch.load(args.model_path, map_location="cpu")
        if "state_dict" in ckpt:  # generate_mask.py:197-200
            ckpt = ckpt["state_dict"]
        eng.load_state_dict(ckpt, strict=False)
    else:
        print("warning: no --model_path, using zero-initialised weights", file=sys.stderr)
    return eng


def _first_n(loader, n, args):
    """torch.utils.data.Subset(dataset, range(n)) + DataLoader(shuffle=False) (main_forget.py:168-171)"""
    from .device_data import DeviceLoader
    if isinstance(loader, DeviceLoader):
        return DeviceLoader(loader.dataset, loader.indices[:n], batch_size=args.batch_size, shuffle=False, augment=False)
    sub = torch.utils.data.Subset(loader.dataset, list(range(min(n, len(loader.dataset)))))
    return torch.utils.data.DataLoader(sub, batch_size=args.batch_size, shuffle=False)


def generate_mask_main(args):
    torch.manual_seed(args.seed)
    os.makedirs(args.save_dir, exist_ok=True)
    loaders = make_loaders(args)
    engine = _engine_from_args(args)
    infos = save_gradient_ratio(OrderedDict(forget=loaders["forget"]), engine, torch.nn.CrossEntropyLoss(), args)
    for r, info in infos.items():
        print(f"with_{r}.pt: threshold |g| = {info.thr_value:.6g}, ties at threshold = {info.n_equal}")


def unlearn_main(args, with_mask: bool):
    torch.manual_seed(args.seed)
    os.makedirs(args.save_dir, exist_ok=True)
    loaders = make_loaders(args)
    engine = _engine_from_args(args)
    mask = None
    if with_mask:
        if not args.mask_path:
            raise SystemExit("main_random needs --mask_path (main_random.py:133-140 raises NameError without it)")
        mask = args.mask_path   # resolved by the engine: the packed side-car <mask_path>.bits when present, else the int64 dict
    method = get_unlearn_method(args.unlearn)
    method(loaders, engine, torch.nn.CrossEntropyLoss(), args, mask) if mask is not None else \
        method(loaders, engine, torch.nn.CrossEntropyLoss(), args)
    # main_forget.py:141-151: every loader in test mode (no augmentation), trainer/val.py validate on the engine's kernels
    crit = torch.nn.CrossEntropyLoss()
    eval_loaders = make_loaders(args, test_mode=True) if getattr(args, "device_data", False) else loaders
    evaluation_result = {"accuracy": {}}
    for name, ld in eval_loaders.items():
        evaluation_result["accuracy"][name] = validate(ld, engine, crit, args)
        print(f"{name} acc: {evaluation_result['accuracy'][name]}")
    if getattr(args, "mia", False):
        # main_forget.py:158-183 forget-efficacy MIA: shadow_train = the first len(test) retain samples, shadow_test = test,
        # target_test = forget
        test_len = len(eval_loaders["test"].dataset) if not getattr(args, "device_data", False) else eval_loaders["test"].indices.numel()
        shadow_train = _first_n(eval_loaders["retain"], test_len, args)
        evaluation_result["SVC_MIA_forget_efficacy"] = SVC_MIA(shadow_train=shadow_train, shadow_test=eval_loaders["test"],
                                                                target_train=None, target_test=eval_loaders["forget"],
                                                                model=engine)
        print("SVC_MIA_forget_efficacy", evaluation_result["SVC_MIA_forget_efficacy"])
    state = {"state_dict": engine.state_dict(), "evaluation_result": evaluation_result}  # impl.py:21-30
    torch.save(state, os.path.join(args.save_dir, str(args.unlearn) + "checkpoint.pth.tar"))  # utils.py:44-52
    torch.save(evaluation_result, os.path.join(args.save_dir, str(args.unlearn) + "eval_result.pth.tar"))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("generate_mask", "main_random", "main_forget"):
        raise SystemExit(__doc__)
    cmd, args = argv[0], arg_parser.parse_args(argv[1:])
    if args.save_dir is None:
        raise SystemExit("--save_dir is required")
    if cmd == "generate_mask":
        generate_mask_main(args)
    else:
        unlearn_main(args, with_mask=(cmd == "main_random"))


if __name__ == "__main__":
    main()
