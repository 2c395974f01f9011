"""Minimal end-to-end run (reference examples/mini_example.py): MNIST-shaped data, 10 clients of which
4 mount the ALIE attack, plain mean aggregation.  No Ray, no cluster: run it as is (CPU / 1 GPU) or under
``torchrun --nproc-per-node N`` for N trainer shards.  ``--real-mnist`` downloads MNIST through
torchvision; the default is the synthetic MNIST-shaped dataset (no network needed)."""
import argparse


from blades_b200 import Simulator
from blades_b200.comm.group import init_world, shutdown
from blades_b200.datasets import MNIST, SyntheticMNIST
from blades_b200.models.mnist import MLP


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--real-mnist", action="store_true")
    ap.add_argument("--rounds", type=int, default=100)
    ap.add_argument("--use-cuda", action="store_true")
    args = ap.parse_args()
    init_world(use_cuda=args.use_cuda)
    ds_cls = MNIST if args.real_mnist else SyntheticMNIST
    mnist = ds_cls(data_root="./data", train_bs=32, num_clients=10, seed=0)
    conf_params = {
        "dataset": mnist,
        "aggregator": "mean",              # defense: robust aggregation
        "num_byzantine": 4,                # number of Byzantine clients
        "attack": "alie",                  # attack strategy
        "attack_kws": {"num_clients": 10, "num_byzantine": 4},
        "num_actors": 4,                   # accepted for API parity (the world size decides)
        "use_cuda": args.use_cuda,
        "seed": 1,                         # reproducibility
    }
    simulator = Simulator(**conf_params)
    run_params = {
        "model": MLP(),                    # global model
        "server_optimizer": "SGD",
        "client_optimizer": "SGD",
        "loss": "crossentropy",
        "global_rounds": args.rounds,
        "local_steps": 50,
        "client_lr": 0.1,
        "server_lr": 1.0,
        "validate_interval": 10,
    }
    times = simulator.run(**run_params)
    print(f"{len(times)} rounds, {sum(times):.2f} s total")
    shutdown()


if __name__ == "__main__":
    main()
