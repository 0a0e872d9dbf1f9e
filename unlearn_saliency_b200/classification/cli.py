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


@torch.no_grad()
def validate(loader, engine) -> float:
    """top-1 accuracy in percent (trainer/val.py:6-72) with the engine's eval-mode forward"""
    correct = total = 0
    for x, y in loader:
        logits = engine.forward(x.to(engine.device, non_blocking=True).float().contiguous())
        correct += int((logits.argmax(1).cpu() == y).sum())
        total += y.numel()
    return 100.0 * correct / max(1, total)


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
        mask = torch.load(args.mask_path, map_location=engine.device)
    method = get_unlearn_method(args.unlearn)
    method(loaders, engine, torch.nn.CrossEntropyLoss(), args, mask) if mask is not None else \
        method(loaders, engine, torch.nn.CrossEntropyLoss(), args)
    evaluation_result = {"accuracy": {name: validate(ld, engine) for name, ld in loaders.items()}}
    for name, acc in evaluation_result["accuracy"].items():
        print(f"{name} acc: {acc:.3f}")
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
