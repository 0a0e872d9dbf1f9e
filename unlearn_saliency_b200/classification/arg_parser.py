"""CLI flags of the Classification scripts that reach the SalUn hot path -- same names and defaults as
Classification/arg_parser.py:4-145 (flags of the pruning / rewinding / imagenet side paths are accepted and ignored or
rejected where the mirror does not serve them)."""
import argparse


def build_parser():
    p = argparse.ArgumentParser(description="SalUn Classification entry points on the sm_100a engine")
    # dataset
    p.add_argument("--data", type=str, default="../data", help="location of the data corpus")
    p.add_argument("--dataset", type=str, default="cifar10", help="dataset")
    p.add_argument("--input_size", type=int, default=32, help="size of input images")
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--num_classes", type=int, default=10)
    # architecture
    p.add_argument("--arch", type=str, default="resnet18", help="model architecture")
    p.add_argument("--imagenet_arch", action="store_true", help="architecture for imagenet size samples")
    # general
    p.add_argument("--seed", default=2, type=int, help="random seed")
    p.add_argument("--train_seed", default=1, type=int)
    p.add_argument("--gpu", type=int, default=0, help="gpu device id")
    p.add_argument("--workers", type=int, default=4)
    p.add_argument("--resume", action="store_true")
    p.add_argument("--checkpoint", type=str, default=None)
    p.add_argument("--save_dir", default=None, type=str, help="The directory used to save the trained models")
    p.add_argument("--model_path", type=str, default=None, help="the path of original model")
    # training
    p.add_argument("--batch_size", type=int, default=256)
    p.add_argument("--lr", default=0.1, type=float)
    p.add_argument("--momentum", default=0.9, type=float)
    p.add_argument("--weight_decay", default=5e-4, type=float)
    p.add_argument("--epochs", default=182, type=int)
    p.add_argument("--warmup", default=0, type=int)
    p.add_argument("--print_freq", default=50, type=int)
    p.add_argument("--decreasing_lr", default="91,136")
    p.add_argument("--no-aug", action="store_true", default=False)
    p.add_argument("--no-l1-epochs", default=0, type=int)
    p.add_argument("--rewind_epoch", default=0, type=int)
    p.add_argument("--rewind_pth", default=None, type=str)
    # unlearn
    p.add_argument("--unlearn", type=str, default="retrain", help="method to unlearn")
    p.add_argument("--unlearn_lr", default=0.01, type=float)
    p.add_argument("--unlearn_epochs", default=10, type=int)
    p.add_argument("--num_indexes_to_replace", type=int, default=None, help="Number of data to forget")
    p.add_argument("--class_to_replace", type=int, default=-1, help="Specific class to forget")
    p.add_argument("--indexes_to_replace", type=list, default=None)
    p.add_argument("--alpha", default=0.2, type=float)
    p.add_argument("--mask_path", default=None, type=str, help="the path of saliency map")
    p.add_argument("--no_l1_epochs", default=0, type=int, help="FT_l1: epochs without the l1 penalty at the end (FT.py:124)")
    # additions of this mirror (no counterpart in the reference)
    p.add_argument("--device_data", action="store_true",
                   help="keep the dataset resident on the GPU as uint8 and run gather + RandomCrop(32,4) + flip + ToTensor as "
                        "one kernel per batch (device_data.py) instead of the host DataLoader")
    p.add_argument("--mia", action="store_true", help="also run the SVC membership-inference evaluation (main_forget.py:158-183)")
    p.add_argument("--precision", type=str, default=None, choices=["bf16", "split"],
                   help="engine build: default split (fp32-class) for generate_mask, bf16 for the unlearning steps")
    p.add_argument("--synthetic", type=int, default=0, metavar="N",
                   help="use N synthetic CIFAR-shaped training images instead of the dataset on disk (offline runs)")
    return p


def parse_args(argv=None):
    return build_parser().parse_args(argv)
