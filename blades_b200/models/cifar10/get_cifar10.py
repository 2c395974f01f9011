"""Standalone CIFAR-10 partition generator (reference models/cifar10/get_cifar10.py): writes the
4-pickle ``data/data_cache[_alpha<a>].obj`` file (train ids, train data, test ids, test data)."""
import argparse
import os
import pickle

import numpy as np

from ...datasets.basedataset import partition


def generate_datasets(iid=False, alpha=1.0, num_clients=100, root="./data", loader=None):
    if loader is None:
        import torchvision
        tr = torchvision.datasets.CIFAR10(train=True, download=True, root=root)
        te = torchvision.datasets.CIFAR10(train=False, download=True, root=root)
        x_tr, y_tr, x_te, y_te = tr.data, np.array(tr.targets), te.data, np.array(te.targets)
    else:
        x_tr, y_tr, x_te, y_te = loader()
    x_tr = np.transpose(x_tr.astype('float32') / 255.0, (0, 3, 1, 2))
    x_te = np.transpose(x_te.astype('float32') / 255.0, (0, 3, 1, 2))
    ids, train, _, test = partition(x_tr, y_tr, x_te, y_te, num_clients, iid, alpha, 1234, 10)
    os.makedirs(root, exist_ok=True)
    path = os.path.join(root, 'data_cache' + ("" if iid else "_alpha" + str(alpha)) + '.obj')
    with open(path, 'wb') as f:
        for obj in (ids, train, ids, test):
            pickle.dump(obj, f)
    return path


if __name__ == '__main__':
    p = argparse.ArgumentParser()
    g = p.add_mutually_exclusive_group(required=False)
    g.add_argument('--iid', dest='iid', action='store_true')
    g.add_argument('--noniid', dest='iid', action='store_false')
    p.add_argument('--alpha', type=float, default=0.1)
    p.add_argument('--num_clients', type=int, default=20)
    p.set_defaults(iid=True)
    a = p.parse_args()
    print(generate_datasets(a.iid, alpha=a.alpha, num_clients=a.num_clients))
