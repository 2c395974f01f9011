"""CLI wrapper around :mod:`blades_b200.models.utils.leaf` (reference models/utils/split_data.py)."""
import argparse
import json
import os

from . import leaf


def create_jsons_for(user_files, which_set, max_users, include_hierarchy, subdir='.', out_dir='.'):
    """Split-by-user packing (reference split_data.py:16-75): ``user_files`` is a list of ``(user, num_samples,
    file)`` -- or ``(user, hierarchy, num_samples, file)`` -- tuples; users are read from ``subdir/file`` and written
    to ``out_dir/<stem>_<which_set>_<k>.json`` in groups of at most ``max_users``.  Returns the written paths."""
    written, users, hier, ns, data, k = [], [], [], [], {}, 0
    for i, t in enumerate(user_files):
        u, h, n, f = t if include_hierarchy else (t[0], None, t[1], t[2])
        with open(os.path.join(subdir, f)) as inf:
            data[u] = json.load(inf)['user_data'][u]
        users.append(u), ns.append(n)
        if include_hierarchy:
            hier.append(h)
        if len(users) == max_users or i == len(user_files) - 1:
            out = {'users': users, 'num_samples': ns, 'user_data': data}
            if include_hierarchy:
                out['hierarchies'] = hier
            path = os.path.join(out_dir, '%s_%s_%d.json' % (os.path.splitext(f)[0], which_set, k))
            os.makedirs(out_dir, exist_ok=True)
            with open(path, 'w') as outf:
                json.dump(out, outf)
            written.append(path)
            users, hier, ns, data, k = [], [], [], {}, k + 1
    return written


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
    tool = 'split_data'
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
