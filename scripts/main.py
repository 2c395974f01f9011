"""Canonical experiment driver (reference scripts/main.py / scripts/cifar10.py): federated dataset,
built-in attack + robust aggregator, CCTNet / MLP / ResNet, MultiStepLR client schedule.

    python scripts/main.py --dataset synthetic-cifar10 --attack alie --agg trimmedmean --global_round 20
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/main.py --use-cuda ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from args import parse_arguments
from blades_b200 import Simulator
from blades_b200 import datasets as D
from blades_b200.comm.group import init_world, shutdown
from blades_b200.models import MLP, CCTNet, resnet18, resnet50


def build_dataset(o):
    kw = dict(data_root=o.data_root, train_bs=o.batch_size, num_clients=o.num_clients, seed=o.seed)
    table = {"mnist": D.MNIST, "cifar10": D.CIFAR10, "cifar100": D.CIFAR100, "synthetic-mnist": D.SyntheticMNIST,
             "synthetic-cifar10": D.SyntheticCIFAR10, "synthetic-cifar100": D.SyntheticCIFAR100}
    return table[o.dataset](**kw)


def build_model(o):
    name = o.model
    if name == "auto":
        name = "mlp" if "mnist" in o.dataset else "cct"
    classes = 100 if "cifar100" in o.dataset else 10
    return {"mlp": lambda: MLP(), "cct": lambda: CCTNet(classes), "resnet18": lambda: resnet18(classes),
            "resnet50": lambda: resnet50(classes)}[name]()


def main(argv=None):
    o = parse_arguments(argv)
    world = init_world(use_cuda=o.use_cuda)
    os.makedirs(o.log_dir, exist_ok=True)
    sim = Simulator(dataset=build_dataset(o), aggregator=o.agg, aggregator_kws=o.agg_args.get(o.agg, {}),
                    num_byzantine=o.num_byzantine if o.attack else 0, attack=o.attack,
                    attack_kws=o.attack_args[o.attack], use_cuda=o.use_cuda, log_path=o.log_dir, seed=o.seed,
                    num_actors=o.num_actors, progress=world.rank == 0)
    model = build_model(o)
    # the reference schedules the client lr by attaching MultiStepLR to a dummy optimizer (cifar10.py:43-46)
    dummy = torch.optim.Adam(model.parameters(), lr=o.lr)
    sched = torch.optim.lr_scheduler.MultiStepLR(dummy, milestones=[150, 300, 500], gamma=0.5)
    times = sim.run(model=model, server_optimizer="SGD", client_optimizer=dummy, loss="crossentropy",
                    global_rounds=o.global_round, local_steps=o.local_round, server_lr=1.0, client_lr=o.lr,
                    validate_interval=o.log_interval, test_batch_size=o.test_batch_size, client_lr_scheduler=sched)
    if world.rank == 0:
        print(f"{len(times)} rounds, mean {sum(times) / max(len(times), 1):.4f} s/round; logs in {o.log_dir}")
    shutdown()


if __name__ == "__main__":
    main()
