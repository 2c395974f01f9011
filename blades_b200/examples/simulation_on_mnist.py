"""Aggregator sweep on (synthetic-)MNIST under the IPM attack -- the working counterpart of the reference's
``examples/Simulation on MNIST.py``: for each defence run the same federation (20 clients, 8 Byzantine, IPM with
epsilon = 100), then read every run's ``stats`` log back and tabulate / plot test accuracy per round.

    python -m blades_b200.examples.simulation_on_mnist [--real-mnist] [--use-cuda] [--rounds 10] [--plot out.png]
"""
import argparse
import ast
import os

from blades_b200 import Simulator
from blades_b200.comm.group import init_world, shutdown
from blades_b200.datasets import MNIST, SyntheticMNIST
from blades_b200.models.mnist import MLP

AGGS = {
    "mean": {},
    "trimmedmean": {"nb": 8},
    "geomed": {},
    "median": {},
    "clippedclustering": {},
}


def read_stats(path: str, kind: str = "test"):
    """The ``stats`` log holds one python-dict literal per line (reference utils.py:67-95 format)."""
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            rec = ast.literal_eval(line.replace("nan", "None").replace("inf", "None"))   # diverged runs log nan/inf
            if rec["_meta"]["type"] == kind:
                rows.append(rec)
    return rows


def main(rounds: int = 10, local_steps: int = 10, real_mnist: bool = False, use_cuda: bool = False,
         out_root: str = "./outputs", plot: str = ""):
    world = init_world(use_cuda=use_cuda)
    ds_cls = MNIST if real_mnist else SyntheticMNIST
    table = []
    for agg, kws in AGGS.items():
        data = ds_cls(data_root="./data", train_bs=32, num_clients=20, seed=1)
        sim = Simulator(dataset=data, aggregator=agg, aggregator_kws=dict(kws), num_byzantine=8, attack="ipm",
                        attack_kws={"epsilon": 100}, num_actors=1, use_cuda=use_cuda, seed=1,
                        log_path=os.path.join(out_root, agg), progress=False)
        sim.run(model=MLP(), server_optimizer="SGD", client_optimizer="SGD", loss="crossentropy",
                global_rounds=rounds, local_steps=local_steps, server_lr=1.0, client_lr=0.1, validate_interval=1)
        if world.rank == 0:
            for rec in read_stats(os.path.join(out_root, agg, "stats")):
                table.append({"Round Number": rec["Round"], "Accuracy (%)": rec["top1"], "Loss": rec["Loss"], "AGG": agg})
    if world.rank == 0:
        import pandas as pd
        df = pd.DataFrame(table)
        last = df[df["Round Number"] == df["Round Number"].max()].set_index("AGG")[["Accuracy (%)", "Loss"]]
        print(last.to_string())
        if plot:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            for agg, g in df.groupby("AGG"):
                plt.plot(g["Round Number"], g["Accuracy (%)"], label=agg)
            plt.xlabel("Round Number")
            plt.ylabel("Accuracy (%)")
            plt.legend()
            plt.savefig(plot, dpi=120)
        shutdown()
        return df
    shutdown()
    return None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--local-steps", type=int, default=10)
    ap.add_argument("--real-mnist", action="store_true")
    ap.add_argument("--use-cuda", action="store_true")
    ap.add_argument("--plot", default="")
    a = ap.parse_args()
    main(a.rounds, a.local_steps, a.real_mnist, a.use_cuda, plot=a.plot)
