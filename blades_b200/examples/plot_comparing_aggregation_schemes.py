"""Toy comparison of the aggregation schemes (reference examples/plot_comparing_aggregation_schemes.py):
60 benign points ~ N((0,0), 20 I) and 40 outliers ~ N((30,30), 60 I) in 2-D through eight aggregators.
Expected: Mean and Clustering are dragged towards the outliers, the others stay with the benign cloud.
Prints a table; draws the figure when matplotlib is installed."""
import numpy as np
import torch

from blades_b200.aggregators import (Autogm, Clippedclustering, Clustering, Geomed, Krum, Mean, Median,
                                     Trimmedmean)


def make_scene(seed=1):
    rng = np.random.RandomState(seed)
    benign = rng.multivariate_normal([0, 0], 20 * np.eye(2), 60)
    outliers = rng.multivariate_normal([30, 30], 60 * np.eye(2), 40)
    return benign, outliers


def run(seed=1):
    benign, outliers = make_scene(seed)
    U = torch.tensor(np.concatenate([benign, outliers]), dtype=torch.float32)
    aggs = {"Mean": Mean(), "Krum": Krum(len(U), len(outliers)), "GeoMed": Geomed(), "Median": Median(),
            "AutoGM": Autogm(lamb=1.0), "TrimmedMean": Trimmedmean(nb=len(outliers)),
            "Clustering": Clustering(), "ClippedClustering": Clippedclustering()}
    return benign, outliers, {k: a(U.clone()).numpy() for k, a in aggs.items()}


def main():
    benign, outliers, res = run()
    for k, v in res.items():
        print(f"{k:18s} ({v[0]:8.3f}, {v[1]:8.3f})")
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return res
    fig, ax = plt.subplots(figsize=(7, 7))
    ax.scatter(*benign.T, s=12, c="tab:blue", label="benign")
    ax.scatter(*outliers.T, s=12, c="tab:red", label="outliers")
    for k, v in res.items():
        ax.scatter(v[0], v[1], marker="*", s=160, label=k)
    ax.legend()
    fig.savefig("aggregation_schemes.png", dpi=120)
    return res


if __name__ == "__main__":
    main()
