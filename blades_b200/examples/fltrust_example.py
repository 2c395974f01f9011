"""FLTrust with one trusted client against ALIE (working version of the reference's ``todo_fltrusted_example.py``,
whose stale ``attack_params`` kwarg raises in ``Simulator.__init__``, simulator.py:84-88).

    python -m blades_b200.examples.fltrust_example [--use-cuda] [--rounds 400]
"""
import argparse

from blades_b200 import Simulator
from blades_b200.comm.group import init_world, shutdown
from blades_b200.datasets import MNIST, SyntheticMNIST
from blades_b200.models.mnist import MLP


def main(rounds: int = 400, real_mnist: bool = False, use_cuda: bool = False, local_steps: int = 2):
    init_world(use_cuda=use_cuda)
    ds_cls = MNIST if real_mnist else SyntheticMNIST
    mnist = ds_cls(data_root="./data", train_bs=32, num_clients=10, seed=1)
    simulator = Simulator(dataset=mnist, aggregator="fltrust", num_byzantine=3, attack="alie",
                          attack_kws={"num_clients": 10, "num_byzantine": 3}, num_actors=4, use_cuda=use_cuda,
                          seed=1, progress=False)
    trusted_id = simulator.get_clients()[-1].id()         # the server's root dataset lives on the last client
    simulator.set_trusted_clients([trusted_id])
    simulator.run(model=MLP(), server_optimizer="SGD", client_optimizer="SGD", loss="crossentropy",
                  global_rounds=rounds, local_steps=local_steps, server_lr=1.0, client_lr=0.1,
                  validate_interval=max(1, rounds // 10))
    shutdown()
    return simulator


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=400)
    ap.add_argument("--real-mnist", action="store_true")
    ap.add_argument("--use-cuda", action="store_true")
    a = ap.parse_args()
    main(a.rounds, a.real_mnist, a.use_cuda)
