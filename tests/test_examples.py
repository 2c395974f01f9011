"""The shipped examples / drivers run (they are the reference's only 'tests')."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_toy_scene_example():
    from blades_b200.examples.plot_comparing_aggregation_schemes import run
    _, _, res = run()
    assert np.linalg.norm(res["Mean"]) > 10
    for k in ("Krum", "GeoMed", "Median", "AutoGM", "TrimmedMean", "ClippedClustering"):
        assert np.linalg.norm(res[k]) < 10, k


def test_customize_attack_example(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from blades_b200.examples.customize_attack import main
    sim = main(rounds=2)
    assert sum(c.is_byzantine() for c in sim.get_clients()) == 3


def test_customize_aggregator_example(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from blades_b200.examples.customize_aggregator import main
    main(rounds=1)


def test_main_script_cli(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "main.py"), "--dataset", "synthetic-mnist",
                          "--attack", "ipm", "--agg", "geomed", "--global_round", "2", "--local_round", "2",
                          "--num_clients", "6", "--num_byzantine", "2", "--log_interval", "1",
                          "--data_root", str(tmp_path / "data")], capture_output=True, text=True, env=env, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "2 rounds" in out.stdout


def test_aggregator_sweep_example(tmp_path, monkeypatch):
    """examples/simulation_on_mnist.py: five defences under IPM, stats logs read back into one table."""
    monkeypatch.chdir(tmp_path)
    from blades_b200.examples.simulation_on_mnist import AGGS, main
    df = main(rounds=2, local_steps=2, out_root=str(tmp_path / "outputs"))
    assert set(df["AGG"]) == set(AGGS) and len(df) == 2 * len(AGGS)
    assert {"Round Number", "Accuracy (%)", "Loss"} <= set(df.columns)


def test_fltrust_example(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from blades_b200.examples.fltrust_example import main
    sim = main(rounds=2)
    assert sum(c.is_trusted() for c in sim.get_clients()) == 1


def test_cifar10_example_small(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from blades_b200.examples.cifar10_example import main
    _, times = main(dataset="cifar10", rounds=1, local_steps=1, num_clients=4, num_byzantine=1)
    assert len(times) == 1
