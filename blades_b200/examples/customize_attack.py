"""Custom attack (reference examples/customize_attack.py): subclass ``ByzantineClient``, override any of
``on_train_batch_begin`` / ``local_training`` / ``omniscient_callback`` and install the clients with
``Simulator.register_attackers``."""
import torch

from blades_b200 import ByzantineClient, Simulator
from blades_b200.datasets import SyntheticMNIST
from blades_b200.models.mnist import MLP


class MaliciousClient(ByzantineClient):
    def __init__(self, scale=10.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.scale = scale

    def on_train_batch_begin(self, data, target, logs=None):
        return data, 9 - target                       # train on flipped labels ...

    def omniscient_callback(self, simulator):
        # ... then, knowing everybody's update, push against the honest mean
        honest = [c.get_update() for c in simulator.get_clients() if not c.is_byzantine()]
        self.save_update(-self.scale * torch.stack(honest).mean(0))


def main(rounds=5):
    dataset = SyntheticMNIST(data_root="./data", train_bs=32, num_clients=10, seed=0)
    simulator = Simulator(dataset=dataset, aggregator="median", seed=1, log_path="./outputs/custom_attack")
    simulator.register_attackers([MaliciousClient(scale=5.0) for _ in range(3)])
    simulator.run(model=MLP(), global_rounds=rounds, local_steps=10, client_lr=0.1, server_lr=1.0,
                  validate_interval=1)
    return simulator


if __name__ == "__main__":
    main()
