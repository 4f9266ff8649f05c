"""CPU tests of the row-slab partition (integer maps must be bit-exact) and of the halo lists over a
real 2-process gloo group (the N>1 host logic; the device side is exercised on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from criteria3d_b200 import SoilFluxes3D
from oracle import ORACLE_LIB
from criteria3d_b200.partition import make_slab, slab_catchment, slab_rows
from criteria3d_b200.synth import Catchment, setup


@pytest.mark.parametrize("rows,world", [(7, 2), (64, 8), (1024, 3), (5, 5)])
def test_slab_rows_cover_the_dem_once(rows, world):
    spans = [slab_rows(rows, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == rows
    for a, b in zip(spans, spans[1:]):
        assert a[1] == b[0]
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("shape,kw", [((13, 9, 3), {}), ((12, 6, 20), {"saturated_bottom": True})], ids=["storm", "c4-recipe"])
@pytest.mark.parametrize("world", [2, 3])
def test_local_graph_equals_global_graph_on_owned_nodes(world, shape, kw):
    """Build the whole catchment and every slab with the CPU oracle: for every OWNED node the link
    slot table, mapped to global ids, the geometry and the initial state (incl. the C4 recipe: 20 soil layers, lower third
    saturated) are identical to the global build."""
    if not ORACLE_LIB.exists():
        pytest.skip("oracle library not built")
    sf = SoilFluxes3D(ORACLE_LIB)
    R, C, L = shape
    cat = Catchment(R, C, L, **kw)
    setup(sf, cat, threads=1)
    g_tab = [sf.link_table(s, 0, cat.n_nodes) for s in range(10)]
    g_meta = sf.node_meta(0, cat.n_nodes)
    g_psi = sf.get_field(3, 0, cat.n_nodes)
    g_H = sf.get_field(4, 0, cat.n_nodes)
    owned_total = 0
    for rank in range(world):
        slab = make_slab(R, C, L, world, rank)
        lc = slab_catchment(slab, **kw)
        setup(sf, lc, threads=1)
        l2g = slab.local_to_global()
        own = slab.owned_mask()
        owned_total += int(own.sum())
        assert slab.n_local == lc.n_nodes and int(own.sum()) == slab.n_owned
        for s in range(10):
            lt, li, ar = sf.link_table(s, 0, lc.n_nodes)
            assert np.array_equal(lt[own], g_tab[s][0][l2g[own]]), (rank, s)
            has = own & (lt != 0)
            assert np.array_equal(l2g[li[has]], g_tab[s][1][l2g[has]].astype(np.int64)), (rank, s)
            assert np.array_equal(ar[has], g_tab[s][2][l2g[has]])
        for a, b in zip(sf.node_meta(0, lc.n_nodes), g_meta):
            assert np.array_equal(a[own], b[l2g[own]])
        assert np.array_equal(sf.get_field(4, 0, lc.n_nodes)[own], g_H[l2g[own]])       # z and psi identical
        assert np.array_equal(sf.get_field(3, 0, lc.n_nodes)[own], g_psi[l2g[own]])
        # every link of an owned node points to an owned node or to a ghost that some peer sends
        peers, send, recv = slab.halo()
        ghosts = set(np.concatenate(recv).tolist()) if recv else set()
        for s in range(10):
            lt, li, _ = sf.link_table(s, 0, lc.n_nodes)
            tgt = li[own & (lt != 0)]
            assert all(own[t] or (t in ghosts) for t in tgt.tolist())
    assert owned_total == cat.n_nodes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_worker(rank, world, port, R, C, L, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slab = make_slab(R, C, L, world, rank)
    l2g = slab.local_to_global()
    own = slab.owned_mask()
    x = np.where(own, np.sin(l2g * 0.37) + l2g, -1.0)          # ghosts start wrong
    peers, send, recv = slab.halo()
    ops, bufs = [], []
    for p, s_idx, r_idx in zip(peers, send, recv):
        sb = torch.from_numpy(np.ascontiguousarray(x[s_idx]))
        rb = torch.empty(len(r_idx), dtype=torch.float64)
        ops += [dist.P2POp(dist.isend, sb, p), dist.P2POp(dist.irecv, rb, p)]
        bufs.append((r_idx, rb))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for r_idx, rb in bufs:
        x[r_idx] = rb.numpy()
    ok = np.array_equal(x, np.sin(l2g * 0.37) + l2g)
    n_owned = torch.tensor([float(own.sum())], dtype=torch.float64)
    dist.all_reduce(n_owned)                                   # what sf3d_ext_set_halo is given as n_global_nodes
    q.put((rank, bool(ok), float(n_owned.item()), slab.n_global))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    R, C, L = 11, 6, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, R, C, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert all(n == ng for _, _, n, ng in res), res


@pytest.mark.parametrize("world", [2, 3])
def test_generic_graph_partition_closure_and_halo(world):
    """Non-grid path: partition a node/link graph by y quantiles; one 'Jacobi-like' neighbour sum computed
    on the local graphs after a halo exchange along the send/recv lists equals the global one."""
    if not ORACLE_LIB.exists():
        pytest.skip("oracle library not built")
    from criteria3d_b200.partition import partition_graph
    sf = SoilFluxes3D(ORACLE_LIB)
    valid = np.ones((9, 7), bool)
    valid[0:2, 0:3] = False
    valid[5, 4] = False
    cat = Catchment(9, 7, 3, valid=valid)
    setup(sf, cat, threads=1)
    n = cat.n_nodes
    lt = np.stack([sf.link_table(s, 0, n)[0] for s in range(10)])
    li = np.stack([sf.link_table(s, 0, n)[1] for s in range(10)])
    surf = sf.node_meta(0, n)[0]
    H = sf.get_field(4, 0, n)
    psi = sf.get_field(3, 0, n)
    yy = np.repeat(np.arange(9)[::-1], 7).reshape(9, 7)[valid].astype(float)      # y of each cell, north = large
    y = np.tile(yy, cat.layers)
    parts = partition_graph(y, surf, lt, li, world)
    assert sorted(np.concatenate([p.owned for p in parts]).tolist()) == list(range(n))
    xg = np.sin(np.arange(n) * 0.61) + (H - psi)
    want = np.zeros(n)
    for s in range(10):
        has = lt[s] != 0
        want[has] += xg[li[s][has]]
    # local vectors: owned values known, ghosts filled through the halo lists
    loc = []
    for p in parts:
        x = np.full(len(p.local_to_global), np.nan)
        g2l = p.global_to_local()
        for g in p.owned:
            x[g2l[int(g)]] = xg[g]
        assert p.n_surface_local == int(surf[p.local_to_global].sum())
        assert np.all(surf[p.local_to_global[: p.n_surface_local]] == 1)            # surface nodes first
        loc.append(x)
    for p in parts:
        for peer, s_idx in zip(p.peers, p.send):
            q = parts[peer]
            r_idx = q.recv[q.peers.index(p.rank)]
            assert len(r_idx) == len(s_idx)
            loc[peer][r_idx] = loc[p.rank][s_idx]
    for p in parts:
        g2l = p.global_to_local()
        assert not np.isnan(loc[p.rank]).any()
        for g in p.owned:
            acc = 0.0
            for s in range(10):
                if lt[s][g] != 0:
                    acc += loc[p.rank][g2l[int(li[s][g])]]
            assert acc == pytest.approx(want[g], rel=1e-14, abs=1e-12)


def test_slab_preconditions_are_checked():
    """the slab maps assume a fully valid raster and at least one row per rank"""
    import numpy as np
    import pytest
    from criteria3d_b200.partition import make_slab, slab_catchment
    with pytest.raises(ValueError):
        make_slab(3, 8, 2, world=4, rank=0)                 # fewer rows than ranks: empty slabs
    with pytest.raises(ValueError):
        make_slab(8, 8, 2, world=2, rank=2)
    slab = make_slab(8, 8, 2, world=2, rank=0)
    valid = np.ones((slab.local_rows, 8), bool); valid[0, 0] = False
    with pytest.raises(ValueError):
        slab_catchment(slab, valid=valid)                   # NODATA cells: node ids follow cell ranks, not row arithmetic
