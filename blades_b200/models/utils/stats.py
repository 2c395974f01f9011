"""CLI wrapper around :mod:`blades_b200.models.utils.leaf` (reference models/utils/stats.py)."""
import argparse
import json
import os

from . import leaf


def load_data(name):
    """``(users, num_samples)`` of every user file under ``<name>/data/all_data`` (reference stats.py:26-48)."""
    ds = leaf.load_dir(os.path.join(name, 'data', 'all_data'))
    return list(ds['users']), list(ds['num_samples'])


def print_dataset_stats(name):
    """Print the summary the reference prints (users, samples, mean / std / skew of samples per user);
    the histogram plot is left out (reference stats.py:51-86)."""
    import numpy as np
    users, num_samples = load_data(name)
    n = np.asarray(num_samples, dtype=np.float64)
    std = n.std()
    skew = float(((n - n.mean()) ** 3).mean() / std ** 3) if std > 0 else 0.0
    print('####################################')
    print('DATASET: %s' % name)
    print('%d users' % len(users))
    print('%d samples (total)' % int(n.sum()))
    print('%.2f samples per user (mean)' % n.mean())
    print('num_samples (std): %.2f' % std)
    print('num_samples (std/mean): %.2f' % (std / n.mean() if n.mean() else 0.0))
    print('num_samples (skewness): %.2f' % skew)


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument('--name', required=True, help='dataset directory containing data/all_data/*.json')
    p.add_argument('--seed', type=int, default=None)
    p.add_argument('--fraction', '--frac', dest='fraction', type=float, default=0.1)
    p.add_argument('--iid', action='store_true')
    p.add_argument('--niid', dest='iid', action='store_false')
    p.add_argument('--u', type=float, default=0.01, help='iid: fraction of users')
    p.add_argument('--by_user', action='store_true')
    p.add_argument('--by_sample', dest='by_user', action='store_false')
    p.add_argument('--min_samples', type=int, default=10)
    a = p.parse_args(argv)
    root = os.path.join(a.name, 'data')
    src = next((os.path.join(root, d) for d in ('sampled_data', 'rem_user_data', 'all_data')
                if os.path.isdir(os.path.join(root, d))), root)
    ds = leaf.load_dir(src)
    tool = 'stats'
    if tool == 'sample':
        leaf.save(leaf.sample(ds, a.fraction, a.iid, a.u, a.seed), os.path.join(root, 'sampled_data', 'data.json'))
    elif tool == 'remove_users':
        leaf.save(leaf.remove_users(ds, a.min_samples), os.path.join(root, 'rem_user_data', 'data.json'))
    elif tool == 'split_data':
        tr, te = leaf.split_data(ds, a.fraction if a.fraction > 0.5 else 0.9, a.by_user, a.seed)
        leaf.save(tr, os.path.join(root, 'train', 'data_train.json'))
        leaf.save(te, os.path.join(root, 'test', 'data_test.json'))
    else:
        print(json.dumps(leaf.stats(ds), indent=1))


if __name__ == '__main__':
    main()
