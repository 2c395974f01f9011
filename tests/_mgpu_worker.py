"""Worker for the multi-GPU tests (launched with torch.distributed.run): runs a short simulation and
dumps the final flat parameter vector + primitive outputs of this rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from blades_b200 import Simulator
from blades_b200.comm.group import init_world, shutdown
from blades_b200.datasets import synthetic_fldataset
from blades_b200.models import MLP, resnet18


def main():
    out_dir, agg, attack, model_name, n_clients = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
    world = init_world(use_cuda=True)
    attack = None if attack == "none" else attack
    f = max(1, n_clients // 5) if attack else 0
    shape = (28, 28) if model_name == "mlp" else (3, 32, 32)
    ds = synthetic_fldataset(n_clients, shape=shape, train_bs=8, train_per_client=16, test_per_client=8, seed=3,
                             separation=2.0)
    akw = {"num_clients": n_clients, "num_byzantine": f} if attack == "alie" else None
    gkw = {"nb": f} if agg == "trimmedmean" else ({"num_clients": n_clients, "num_byzantine": f} if agg == "krum" else None)
    sim = Simulator(ds, num_byzantine=f, attack=attack, attack_kws=akw, aggregator=agg, aggregator_kws=gkw,
                    use_cuda=True, seed=1, log_path=os.path.join(out_dir, f"logs_{world.size}"), progress=False)
    torch.manual_seed(5)
    m = MLP() if model_name == "mlp" else resnet18(10)
    sim.run(m, global_rounds=4, local_steps=1, server_lr=1.0, client_lr=0.05, validate_interval=4)
    torch.cuda.synchronize()
    vec = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu()
    torch.save(vec, os.path.join(out_dir, f"theta_{agg}_{attack}_{model_name}_{world.size}_{world.rank}.pt"))
    shutdown()


if __name__ == "__main__":
    try:
        main()
    except Exception:
        import traceback
        with open(os.path.join(sys.argv[1], f"error_rank{os.environ.get('RANK', '0')}.txt"), "w") as f:
            traceback.print_exc(file=f)
        traceback.print_exc()
        raise
