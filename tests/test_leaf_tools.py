import numpy as np

from blades_b200.models.utils import leaf
from blades_b200.models.utils.util import iid_divide


def _toy(n_users=12, seed=0):
    rng = np.random.default_rng(seed)
    users = [f"u{i}" for i in range(n_users)]
    data = {}
    for i, u in enumerate(users):
        k = 4 + 3 * i
        data[u] = {"x": rng.standard_normal((k, 5)).tolist(), "y": rng.integers(0, 3, k).tolist()}
    return {"users": users, "num_samples": [len(data[u]["y"]) for u in users], "user_data": data}


def test_leaf_pipeline(tmp_path):
    ds = _toy()
    assert [len(g) for g in iid_divide(list(range(10)), 4)] == [3, 3, 2, 2]
    kept = leaf.remove_users(ds, min_samples=10)
    assert all(n >= 10 for n in kept["num_samples"]) and len(kept["users"]) == 10
    s = leaf.sample(ds, fraction=0.3, iid=False, seed=1)
    assert 0 < sum(s["num_samples"]) <= sum(ds["num_samples"])
    s_iid = leaf.sample(ds, fraction=0.5, iid=True, user_fraction=0.25, seed=1)
    assert len(s_iid["users"]) == 3 and abs(max(s_iid["num_samples"]) - min(s_iid["num_samples"])) <= 1
    tr, te = leaf.split_data(ds, frac=0.75, by_user=False, seed=2)
    assert tr["users"] == te["users"]
    assert all(a + b == n for a, b, n in zip(tr["num_samples"], te["num_samples"], ds["num_samples"]))
    tr_u, te_u = leaf.split_data(ds, frac=0.75, by_user=True, seed=2)
    assert set(tr_u["users"]).isdisjoint(te_u["users"]) and len(tr_u["users"]) == 9
    st = leaf.stats(ds)
    assert st["users"] == 12 and st["samples"] == sum(ds["num_samples"])
    leaf.save(tr, str(tmp_path / "d" / "a.json"))
    assert leaf.load_dir(str(tmp_path / "d"))["users"] == tr["users"]
    fl = leaf.to_fldataset(tr, te, train_bs=4)
    x, y = fl.get_train_data(0, 1)[0]
    assert x.shape[1] == 5 and len(y) == x.shape[0]


def test_get_cifar10_cache(tmp_path):
    import pickle
    from blades_b200.models.cifar10.get_cifar10 import generate_datasets
    rng = np.random.default_rng(0)
    loader = lambda: (rng.integers(0, 255, (200, 32, 32, 3), dtype=np.uint8), rng.integers(0, 10, 200),
                      rng.integers(0, 255, (40, 32, 32, 3), dtype=np.uint8), rng.integers(0, 10, 40))
    path = generate_datasets(iid=True, num_clients=4, root=str(tmp_path), loader=loader)
    with open(path, "rb") as f:
        ids, train, ids2, test = [pickle.load(f) for _ in range(4)]
    assert ids == ["0", "1", "2", "3"] and train["0"]["x"].shape == (50, 3, 32, 32) and test["3"]["y"].shape == (10,)


def test_reference_helper_names(tmp_path, capsys):
    """``stats.load_data / print_dataset_stats`` and ``split_data.create_jsons_for`` (names of the reference scripts)."""
    import json
    from blades_b200.models.utils import split_data, stats
    root = tmp_path / "ds" / "data" / "all_data"
    root.mkdir(parents=True)
    users = {"u%d" % i: {"x": [[float(i)]] * (3 + i), "y": [i % 2] * (3 + i)} for i in range(5)}
    (root / "all_data_0.json").write_text(json.dumps({"users": list(users), "num_samples": [3 + i for i in range(5)],
                                                      "user_data": users}))
    us, ns = stats.load_data(str(tmp_path / "ds"))
    assert sorted(us) == sorted(users) and sum(ns) == sum(3 + i for i in range(5))
    stats.print_dataset_stats(str(tmp_path / "ds"))
    assert "5 users" in capsys.readouterr().out
    files = [(u, 3 + i, "all_data_0.json") for i, u in enumerate(users)]
    out = split_data.create_jsons_for(files, "train", 2, False, subdir=str(root), out_dir=str(tmp_path / "out"))
    assert len(out) == 3
    first = json.loads(open(out[0]).read())
    assert first["users"] == ["u0", "u1"] and first["num_samples"] == [3, 4]
