"""Small helpers kept for parity with the reference's models/utils/util.py."""
import pickle


def save_obj(obj, name):
    with open(name + '.pkl', 'wb') as f:
        pickle.dump(obj, f, pickle.HIGHEST_PROTOCOL)


def load_obj(name):
    with open(name + '.pkl', 'rb') as f:
        return pickle.load(f)


def iid_divide(l, g):
    """Split list ``l`` into ``g`` contiguous groups whose sizes differ by at most one
    (larger groups first) -- same contract as LEAF's helper."""
    n, big = len(l) // g, len(l) % g
    out, pos = [], 0
    for i in range(g):
        k = n + (1 if i < big else 0)
        out.append(l[pos:pos + k])
        pos += k
    return out
