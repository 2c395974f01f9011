"""Custom defense (reference examples/todo_customize_aggregator.py, made runnable): any callable
``inputs -> Tensor[d]`` works; subclassing ``_BaseAggregator`` and using the matrix primitives makes it
run as fused kernels on (sharded) device memory."""
import torch

from blades_b200 import Simulator
from blades_b200.aggregators.base import _BaseAggregator
from blades_b200.datasets import SyntheticMNIST
from blades_b200.models.mnist import MLP


def clipped_mean(clients, max_norm: float = 1.0):
    """Plain-callable form: receives the list of clients (reference convention)."""
    ups = torch.stack([c.get_update() for c in clients])
    scale = torch.clamp(max_norm / ups.norm(dim=1, keepdim=True), max=1.0)
    return (ups * scale).mean(0)


class NormClippedMean(_BaseAggregator):
    """Primitive form: one Gram pass for the norms + one weighted row-combine."""

    def __init__(self, max_norm: float = 1.0):
        super().__init__()
        self.max_norm = max_norm

    def aggregate(self, matrix):
        import numpy as np
        norms = np.sqrt(np.maximum(np.diag(matrix.gram()), 1e-30))
        w = np.minimum(1.0, self.max_norm / norms) / matrix.n_rows
        return matrix.combine(w)


def main(rounds=3):
    for agg in (clipped_mean, NormClippedMean(1.0)):
        dataset = SyntheticMNIST(data_root="./data", train_bs=32, num_clients=10, seed=0)
        sim = Simulator(dataset=dataset, aggregator=agg, num_byzantine=2, attack="signflipping", seed=1,
                        log_path="./outputs/custom_agg")
        sim.run(model=MLP(), global_rounds=rounds, local_steps=5, validate_interval=1)


if __name__ == "__main__":
    main()
