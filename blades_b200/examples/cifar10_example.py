"""CIFAR-10 + CCTNet under attack (working version of the reference's ``todo_cifar10_cpu.py`` / ``todo_cifar10_gpu.py``
and of ``todo_mnist_example.py`` with ``--dataset mnist``).  One process per GPU instead of Ray actors:

    python -m blades_b200.examples.cifar10_example                       # CPU, synthetic CIFAR-10-shaped data
    python -m blades_b200.examples.cifar10_example --use-cuda            # one B200
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 \
        -m blades_b200.examples.cifar10_example --use-cuda               # four trainer shards
"""
import argparse

from blades_b200 import Simulator
from blades_b200.comm.group import init_world, shutdown
from blades_b200.datasets import CIFAR10, MNIST, SyntheticCIFAR10, SyntheticMNIST
from blades_b200.models.cifar10 import CCTNet
from blades_b200.models.mnist import MLP


def main(dataset: str = "cifar10", real_data: bool = False, use_cuda: bool = False, rounds: int = 400,
         local_steps: int = 50, attack: str = "alie", aggregator: str = "clippedclustering", num_clients: int = 20,
         num_byzantine: int = 5):
    init_world(use_cuda=use_cuda)
    if dataset == "mnist":
        data = (MNIST if real_data else SyntheticMNIST)(data_root="./data", train_bs=32, num_clients=num_clients, seed=1)
        model = MLP()
    else:
        data = (CIFAR10 if real_data else SyntheticCIFAR10)(data_root="./data", train_bs=64, num_clients=num_clients,
                                                           seed=1)
        model = CCTNet()
    attack_kws = {"num_clients": num_clients, "num_byzantine": num_byzantine} if attack == "alie" else {}
    sim = Simulator(dataset=data, aggregator=aggregator, num_byzantine=num_byzantine, attack=attack,
                    attack_kws=attack_kws, num_actors=num_clients, use_cuda=use_cuda, seed=1, progress=False)
    times = sim.run(model=model, server_optimizer="SGD", client_optimizer="SGD", loss="crossentropy",
                    global_rounds=rounds, local_steps=local_steps, server_lr=1.0, client_lr=0.1,
                    validate_interval=max(1, rounds // 20))
    shutdown()
    return sim, times


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="cifar10", choices=["cifar10", "mnist"])
    ap.add_argument("--real-data", action="store_true", help="download through torchvision instead of synthetic data")
    ap.add_argument("--use-cuda", action="store_true")
    ap.add_argument("--rounds", type=int, default=400)
    ap.add_argument("--local-steps", type=int, default=50)
    ap.add_argument("--attack", default="alie")
    ap.add_argument("--agg", default="clippedclustering")
    a = ap.parse_args()
    main(a.dataset, a.real_data, a.use_cuda, a.rounds, a.local_steps, a.attack, a.agg)
