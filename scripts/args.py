"""Command-line options of the experiment drivers (flag names follow the reference's
scripts/args.py:7-68 so sweep scripts port unchanged; Ray-only knobs are accepted and ignored)."""
import argparse
import os

import torch

ATTACK_ARGS = {
    None: {}, "none": {}, "noise": {"mean": 0.1, "std": 0.1}, "labelflipping": {}, "signflipping": {},
    "ipm": {"epsilon": 0.5}, "alie": None,        # alie kwargs depend on the client counts (filled below)
}


def agg_args(options):
    return {
        "mean": {}, "median": {}, "geomed": {}, "autogm": {}, "clustering": {}, "clippedclustering": {},
        "centeredclipping": {}, "fltrust": {},
        "trimmedmean": {"nb": options.num_byzantine},
        "krum": {"num_clients": options.num_clients, "num_byzantine": options.num_byzantine},
        "multikrum": {"num_byzantine": options.num_byzantine},
    }


def parse_arguments(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--use-cuda", action="store_true", default=False)
    p.add_argument("--use_actor", action="store_true", default=False, help="(Ray) ignored")
    p.add_argument("--seed", type=int, default=1)
    p.add_argument("--global_round", type=int, default=400)
    p.add_argument("--local_round", type=int, default=50)
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--test_batch_size", type=int, default=128)
    p.add_argument("--log_interval", type=int, default=10)
    p.add_argument("--metrics_name", type=str, default="none")
    p.add_argument("--attack", type=str, default="signflipping")
    p.add_argument("--dataset", type=str, default="cifar10",
                   help="mnist | cifar10 | cifar100 | synthetic-mnist | synthetic-cifar10 | synthetic-cifar100")
    p.add_argument("--agg", type=str, default="clippedclustering")
    p.add_argument("--lr", type=float, default=0.1)
    p.add_argument("--num_actors", type=int, default=20, help="(Ray) ignored: the world size decides")
    p.add_argument("--num_clients", type=int, default=20)
    p.add_argument("--num_byzantine", type=int, default=8)
    p.add_argument("--num_gpus", type=int, default=4, help="(Ray) ignored: launch with torchrun")
    p.add_argument("--model", type=str, default="auto", help="auto | mlp | cct | resnet18 | resnet50")
    p.add_argument("--data_root", type=str, default="./data")
    options = p.parse_args(argv)

    root = os.path.dirname(os.path.abspath(__file__))
    exp_dir = os.path.join(root, f"outputs/{options.dataset}")
    attack = None if options.attack in (None, "none") else options.attack
    options.attack = attack
    akw = ATTACK_ARGS.get(attack, {})
    if attack == "alie":
        akw = {"num_clients": options.num_clients, "num_byzantine": options.num_byzantine}
    options.attack_args = {attack: akw}
    options.agg_args = agg_args(options)

    def tag(d):
        return ("_" + "_".join(f"{k}{v}" for k, v in d.items())) if d else ""
    options.log_dir = (exp_dir + f"/b{options.num_byzantine}_{attack}{tag(akw if attack != 'alie' else {})}"
                       f"_{options.agg}{tag(options.agg_args.get(options.agg, {}))}"
                       f"_lr{options.lr}_bz{options.batch_size}_seed{options.seed}")
    options.use_cuda = torch.cuda.is_available()
    options.gpu_per_actor = 0
    return options


if __name__ != "__main__":
    import sys
    options = parse_arguments(sys.argv[1:] if "pytest" not in sys.modules else [])
