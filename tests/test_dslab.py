"""D-slab mode, host side (CPU, gloo): slab ranges, the exchange regions reported by the library and the
all-reduce + halo swap between layers over a world_size-2 and -3 gloo group."""
import os
import socket

import pytest

from mvsnet_b200 import dslab

SHAPE = dict(n_views=3, depth_num=48, hf=16, wf=24, channels=32, base_filter=8)


def test_slab_ranges():
    assert dslab.slab_range(256, 3, 8) == (96, 128)
    assert [dslab.slab_range(48, r, 3) for r in range(3)] == [(0, 16), (16, 32), (32, 48)]
    with pytest.raises(ValueError):
        dslab.slab_range(48, 0, 4)          # 12 planes per slab: not a multiple of 8
    with pytest.raises(ValueError):
        dslab.slab_range(50, 0, 2)


@pytest.mark.parametrize("world", [2, 3])
def test_regions_are_disjoint_and_inside_the_workspace(world):
    from mvsnet_b200 import _lib as L
    s = SHAPE
    total = L.load().mvsb200_slab_workspace_bytes(s["n_views"], s["depth_num"], world, s["hf"], s["wf"], s["channels"],
                                                  s["base_filter"])
    assert total > 0
    spans = []
    for layer in range(dslab.N_LAYERS):
        r = dslab.layer_regions(layer, s["n_views"], s["depth_num"], world, s["hf"], s["wf"], s["channels"],
                                s["base_filter"])
        off, n = r["stats"]
        assert (n == 0) == (layer == dslab.N_LAYERS - 1)
        if n:
            spans.append((off, off + n))
        for t in r["tensors"]:
            dl = (t["after"] - t["before"]) // t["plane"] - 1          # local planes of this tensor
            assert dl >= 1 and t["first"] == t["before"] + t["plane"] and t["last"] == t["before"] + dl * t["plane"]
            spans.append((t["before"], t["after"] + t["plane"]))
        assert len(r["tensors"]) == (0 if layer == dslab.N_LAYERS - 1 else (2 if layer in (0, 1) else 1))
    fo, fn = dslab.layer_regions(10, s["n_views"], s["depth_num"], world, s["hf"], s["wf"], s["channels"],
                                 s["base_filter"])["filtered"]
    assert fn == s["depth_num"] // world * s["hf"] * s["wf"] * 4
    spans.append((fo, fo + fn))
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0, "exchange regions overlap"
    assert spans[-1][1] <= total


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mvsnet_b200 import _lib as L
        s = SHAPE
        total = L.load().mvsb200_slab_workspace_bytes(s["n_views"], s["depth_num"], world, s["hf"], s["wf"],
                                                      s["channels"], s["base_filter"])
        ok = True
        for layer in (0, 3, 9):                       # two-tensor output, full-resolution conv, transposed conv
            r = dslab.layer_regions(layer, s["n_views"], s["depth_num"], world, s["hf"], s["wf"], s["channels"],
                                    s["base_filter"])
            ws = torch.zeros(total, dtype=torch.uint8)
            off, n = r["stats"]
            ws[off:off + n].view(torch.float64)[:] = float(rank + 1)
            for t in r["tensors"]:
                ws[t["first"]:t["first"] + t["plane"]] = 10 + rank          # boundary planes tagged by owner
                ws[t["last"]:t["last"] + t["plane"]] = 100 + rank
            dslab.exchange_layer(ws, r, rank, world)
            ok &= bool((ws[off:off + n].view(torch.float64) == world * (world + 1) / 2).all())
            for t in r["tensors"]:
                before = ws[t["before"]:t["before"] + t["plane"]]
                after = ws[t["after"]:t["after"] + t["plane"]]
                ok &= bool((before == (100 + rank - 1 if rank > 0 else 0)).all())       # previous rank's last plane
                ok &= bool((after == (10 + rank + 1 if rank < world - 1 else 0)).all())  # next rank's first plane
        ret.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert got == [(r, True) for r in range(world)]


def test_slab_schedule_respects_the_data_flow():
    """Every layer runs after its producers; the schedule covers each layer once."""
    assert sorted(dslab.SLAB_ORDER) == list(range(dslab.N_LAYERS))
    done = set()
    for layer in dslab.SLAB_ORDER:
        assert all(src in done for src in dslab.LAYER_INPUTS[layer]), layer
        done.add(layer)
    # the table mirrors the reference graph: skip adds feed 3dconv5_0 / 6_0 / 6_2 (mvsnetworks.py:148-157)
    assert dslab.LAYER_INPUTS[8] == [7, 5] and dslab.LAYER_INPUTS[9] == [8, 4] and dslab.LAYER_INPUTS[10] == [9, 3]
