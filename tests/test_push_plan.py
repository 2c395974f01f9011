"""Host logic of the push-mode sharded aggregation (no GPU): the copy plan of every rank, taken together, delivers every
needed (row, coordinate) exactly once to the rank that aggregates that coordinate; windows tile the vector."""
import numpy as np
import pytest

from blades_b200.comm.symm import coordinate_shards, round_up
from blades_b200.parallel.sharded import push_plan


@pytest.mark.parametrize("world,n,d,skip", [(2, 10, 5000, set()), (4, 13, 70001, {0, 1, 2}), (8, 100, 300000, set(range(20))),
                                             (8, 9, 1000, {4})])
def test_push_plans_cover_every_needed_element_once(world, n, d, skip):
    sizes = [len(a) for a in np.array_split(np.arange(n), world)]
    row0 = [sum(sizes[:r]) for r in range(world)]
    ld = round_up(d, 64)
    recv_ld = round_up((d + world - 1) // world + 128 * 17, 64)
    # three windows like the ResNet schedule (suffixes of the flat vector, 128-aligned starts), then the whole vector
    cuts = sorted({0, (d // 4 + 127) // 128 * 128, (d // 2 + 127) // 128 * 128, d})
    windows = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)][::-1] + [None]
    for win in windows:
        if win is None:
            shards = coordinate_shards(d, world)
        else:
            lo, hi = win
            shards = [(lo + a, lo + b) for a, b in coordinate_shards(hi - lo, world)]
        # shards tile the window
        assert shards[0][0] == (0 if win is None else win[0]) and shards[-1][1] == (d if win is None else win[1])
        assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
        recv = [np.zeros((n, recv_ld), dtype=np.int32) for _ in range(world)]     # write counters of every landing zone
        src_row = [np.full((n, recv_ld), -1, dtype=np.int64) for _ in range(world)]
        for me in range(world):
            for (g, grow, gcol, src_off, width, height) in push_plan(me, row0, sizes[me], ld, recv_ld, shards, 256, skip):
                assert g != me and width == shards[g][1] - shards[g][0]
                a, c0 = divmod(src_off, ld)
                assert c0 == shards[g][0] and row0[me] + a == grow and a + height <= sizes[me]
                recv[g][grow: grow + height, gcol: gcol + width] += 1
                src_row[g][grow: grow + height, gcol: gcol + width] = np.arange(c0, c0 + width)[None, :]
        for g in range(world):
            d0, d1 = shards[g]
            for i in range(n):
                owner = max(r for r in range(world) if row0[r] <= i)
                want = 0 if (owner == g or i in skip) else 1
                assert (recv[g][i, 256: 256 + (d1 - d0)] == want).all(), (g, i)
                assert recv[g][i].sum() == want * (d1 - d0)
                if want:
                    assert (src_row[g][i, 256: 256 + (d1 - d0)] == np.arange(d0, d1)).all()


def test_window_spans_fit_the_landing_zone():
    d, world = 11181642, 8
    recv_ld = round_up((d + world - 1) // world + 128 * 17, 64)
    col = 0
    for lo, hi in [(2782848, d), (683008, 2782848), (0, 683008)]:
        span = max(b - a for a, b in coordinate_shards(hi - lo, world))
        col += (span + 127) // 128 * 128
    assert col <= recv_ld


def test_pipeline_windows_are_contiguous_rounded_up_and_bounded(monkeypatch):
    """_AggPipeline bookkeeping without a GPU: windows tile [0, d) from the top, start on 128-float boundaries rounded
    UP (coordinates below a reported offset are not final yet), tiny windows are merged into the next one."""
    import types
    from blades_b200.engine import round as R
    d = 11181642
    eng = types.SimpleNamespace(d=d, symm=types.SimpleNamespace(window_span=lambda lo, hi: (hi - lo + 7) // 8), device="cpu")
    launched = []
    monkeypatch.setattr(R._AggPipeline, "__init__", lambda self, e, fn: self.__dict__.update(
        eng=e, fn=fn, hi=e.d, k=0, windows=[], out=None, agg=None, col=0))

    def fake_launch(self, lo, hi, last):
        launched.append((lo, hi, last, self.col))
        self.windows.append((lo, hi))
        self.k += 1
        self.hi = lo
        self.col += (self.eng.symm.window_span(lo, hi) + 127) // 128 * 128
    monkeypatch.setattr(R._AggPipeline, "_launch", fake_launch)
    pipe = R._AggPipeline(eng, None)
    for lo in (2782890, 683009, 157001, 9500, 0):       # layer4, layer3, layer2 (small), layer1 (small), stem
        pipe.progress(lo)
    monkeypatch.setattr("torch.cuda.current_stream", lambda *_: types.SimpleNamespace(wait_stream=lambda s: None))
    eng._agg_stream = None
    pipe.finish()
    # layer2 (4.7 % of d) is just above the 4 % threshold and gets its own window; layer1 (1.3 %) joins the stem
    assert [w[:2] for w in launched] == [(2782976, d), (683136, 2782976), (157056, 683136), (0, 157056)]
    assert [w[2] for w in launched] == [False, False, False, True]
    assert all(w[0] % 128 == 0 for w in launched)
    cols = [w[3] for w in launched]
    assert cols == sorted(cols) and cols[0] == 0


def test_checkpoint_plain_roundtrip_is_weights_only_loadable(tmp_path):
    """Format-3 checkpoints hold tensors and plain containers only: numpy RNG states / cursors survive the conversion
    and the file loads with ``weights_only=True`` (no code execution on resume)."""
    import random
    import torch
    from blades_b200 import checkpoint as ck
    rng = {"numpy": np.random.get_state(), "python": random.getstate(), "torch": torch.get_rng_state()}
    cur = {"client_3": {"perm": np.arange(7)[::-1].copy(), "pos": np.int64(4), "epoch": 2, "f": np.float32(0.5)}}
    payload = {"rng": ck._to_plain([rng]), "data_cursors": ck._to_plain(cur)}
    path = tmp_path / "c.pt"
    torch.save(payload, path)
    back = torch.load(path, map_location="cpu", weights_only=True)
    rng2 = ck._from_plain(back["rng"])[0]
    cur2 = ck._from_plain(back["data_cursors"])
    assert rng2["numpy"][0] == rng["numpy"][0] and (rng2["numpy"][1] == rng["numpy"][1]).all()
    assert rng2["numpy"][1].dtype == np.uint32 and rng2["numpy"][2:] == rng["numpy"][2:]
    assert rng2["python"] == rng["python"] and torch.equal(rng2["torch"], rng["torch"])
    assert (cur2["client_3"]["perm"] == cur["client_3"]["perm"]).all() and cur2["client_3"]["pos"] == 4
    np.random.set_state(rng2["numpy"])
    random.setstate(rng2["python"])


def test_windowed_primitives_tile_to_the_whole_result_on_the_cpu_oracle():
    """The coordinate-wise primitives of a matrix object restricted to a window write only that window of the shared
    result vector; windows that tile [0, d) reproduce the unwindowed result (the contract the pipelined GPU rounds
    rely on), with and without virtual attack rows."""
    import torch
    from blades_b200.parallel.matrix import LocalMatrix, VirtualRows
    torch.manual_seed(0)
    n, d = 12, 1000
    data = torch.randn(n, d)
    windows = [(640, d), (256, 640), (0, 256)]
    for virt in (None, VirtualRows("alie", 0.5, [0, 1, 2]), VirtualRows("ipm", 2.0, [0, 1])):
        for op in (lambda m: m.trimmed_mean(3), lambda m: m.median(), lambda m: m.mean(),
                   lambda m: m.combine(torch.linspace(0.0, 1.0, n))):
            whole = op(LocalMatrix(data.clone(), virtual=virt))
            out = None
            for k, win in enumerate(windows):
                m = LocalMatrix(data.clone(), virtual=virt)
                m.window, m.chunk, m.last_chunk, m.out_buffer = win, k, k == len(windows) - 1, out
                out = op(m)
            assert torch.allclose(out, whole, atol=0, rtol=0)
