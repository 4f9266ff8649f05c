"""Row-slab partition of a DEM catchment over the GPUs of one box.

All layers of a cell column live on the same rank (vertical links never cross); lateral links reach
at most one DEM row, so each rank holds its owned rows plus ONE ghost row per neighbouring slab.
The local raster (ghost rows included) is an ordinary catchment for the library; this module only
produces the integer maps: owned row ranges, local<->global node ids, and the halo send/recv lists
(bit-exact index work, tested on CPU with gloo in tests/test_partition.py)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .synth import Catchment


def slab_rows(rows: int, world: int, rank: int) -> tuple[int, int]:
    """Owned DEM rows [r0, r1) of `rank`: contiguous, sizes differ by at most one row."""
    base, extra = divmod(rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


@dataclass
class Slab:
    rank: int
    world: int
    rows: int                 # global DEM rows
    cols: int
    layers: int               # surface + soil layers
    r0: int                   # first owned global row
    r1: int                   # one past the last owned global row
    top_ghost: int            # 1 if a ghost row precedes the owned rows
    bottom_ghost: int

    @property
    def local_row0(self) -> int:
        return self.r0 - self.top_ghost

    @property
    def local_rows(self) -> int:
        return (self.r1 - self.r0) + self.top_ghost + self.bottom_ghost

    @property
    def n_local(self) -> int:
        return self.layers * self.local_rows * self.cols

    @property
    def n_owned(self) -> int:
        return self.layers * (self.r1 - self.r0) * self.cols

    @property
    def n_global(self) -> int:
        return self.layers * self.rows * self.cols

    # ---- integer maps ------------------------------------------------------------------
    def local_row_nodes(self, local_row: int) -> np.ndarray:
        """local node ids of one local DEM row, all layers, layer-major"""
        per_layer = self.local_rows * self.cols
        base = local_row * self.cols + np.arange(self.cols, dtype=np.int64)
        return (np.arange(self.layers, dtype=np.int64)[:, None] * per_layer + base[None, :]).reshape(-1).astype(np.uint32)

    def local_to_global(self) -> np.ndarray:
        """global node id of every local node (ghosts included)"""
        lay, row, col = np.meshgrid(np.arange(self.layers, dtype=np.int64),
                                    np.arange(self.local_rows, dtype=np.int64) + self.local_row0,
                                    np.arange(self.cols, dtype=np.int64), indexing="ij")
        return (lay * (self.rows * self.cols) + row * self.cols + col).reshape(-1)

    def owned_mask(self) -> np.ndarray:
        m = np.zeros((self.layers, self.local_rows, self.cols), bool)
        m[:, self.top_ghost: self.top_ghost + (self.r1 - self.r0), :] = True
        return m.reshape(-1)

    def halo(self):
        """(peers, send_lists, recv_lists): for each neighbouring rank the owned boundary row to send
        and the ghost row to receive, as local node ids (all layers)."""
        peers, send, recv = [], [], []
        if self.top_ghost:
            peers.append(self.rank - 1)
            send.append(self.local_row_nodes(self.top_ghost))            # first owned row
            recv.append(self.local_row_nodes(0))                          # ghost row above
        if self.bottom_ghost:
            peers.append(self.rank + 1)
            send.append(self.local_row_nodes(self.local_rows - 1 - self.bottom_ghost))   # last owned row
            recv.append(self.local_row_nodes(self.local_rows - 1))       # ghost row below
        return peers, send, recv


def make_slab(rows: int, cols: int, n_soil_layers: int, world: int, rank: int) -> Slab:
    """Row slab of a FULLY VALID rectangular raster (node id = layer * rows * cols + row * cols + col).
    Rasters with NODATA cells number their nodes by cell rank instead: use partition_graph for those."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    if rows < world:
        raise ValueError(f"{rows} DEM rows cannot be split into {world} non-empty row slabs")
    r0, r1 = slab_rows(rows, world, rank)
    return Slab(rank=rank, world=world, rows=rows, cols=cols, layers=n_soil_layers + 1, r0=r0, r1=r1,
                top_ghost=1 if rank > 0 else 0, bottom_ghost=1 if rank < world - 1 else 0)


def slab_catchment(slab: Slab, **kw) -> Catchment:
    """The local raster of a slab: the same seeded generator evaluated on the slab's global rows."""
    if kw.get("valid") is not None:
        raise ValueError("row slabs assume a raster without NODATA cells (the halo lists are computed from "
                         "row / column arithmetic); partition ragged catchments with partition_graph")
    return Catchment(slab.local_rows, slab.cols, slab.layers - 1, row0=slab.local_row0, global_rows=slab.rows, **kw)


# ---------------------------------------------------------------------------------------------
# generic node/link graphs (not rasters): partition by y quantiles, ghost set = link closure
# ---------------------------------------------------------------------------------------------
@dataclass
class RankGraph:
    rank: int
    owned: np.ndarray            # global ids owned by this rank
    ghosts: np.ndarray           # global ids of the halo copies this rank needs
    local_to_global: np.ndarray  # local numbering: owned surface, ghost surface, owned soil, ghost soil
    n_surface_local: int
    peers: list                  # neighbouring ranks
    send: list                   # per peer: LOCAL ids of owned nodes the peer needs (ascending global id)
    recv: list                   # per peer: LOCAL ids of the ghosts owned by that peer (ascending global id)

    def global_to_local(self) -> dict:
        return {int(g): k for k, g in enumerate(self.local_to_global)}


def partition_graph(y: np.ndarray, surface_flag: np.ndarray, link_type: np.ndarray, link_index: np.ndarray,
                    world: int) -> list[RankGraph]:
    """Partition an arbitrary soilFluxes3D graph over `world` ranks (non-grid case).

    link_type / link_index: (10, N) slot tables as the API stores them (slot 0 Up, 1 Down, 2.. Lateral).
    Vertical links never cross ranks: nodes are grouped into columns by following Up links, columns are
    sorted by the y of their top node (north to south, ties by id) and cut into `world` contiguous groups
    of near-equal node count.  Ghosts of a rank = link targets of its owned nodes that another rank owns.
    """
    n = y.shape[0]
    up_t, up_i = link_type[0], link_index[0]
    root = np.arange(n, dtype=np.int64)
    for i in range(n):                       # Up neighbours are created before their lower nodes in every caller
        if up_t[i] != 0:
            j = int(up_i[i])
            root[i] = root[j] if j < i else j
    for _ in range(64):                      # path compression for graphs numbered differently
        nxt = root[root]
        if np.array_equal(nxt, root):
            break
        root = nxt
    cols, inv, counts = np.unique(root, return_inverse=True, return_counts=True)
    order = np.lexsort((cols, -y[cols]))     # north (large y) first, ties by id
    cum = np.cumsum(counts[order])
    owner_of_col = np.empty(len(cols), np.int64)
    owner_of_col[order] = np.minimum((cum - 1) * world // cum[-1], world - 1)
    owner = owner_of_col[inv]

    out = []
    surface_flag = surface_flag.astype(bool)
    for r in range(world):
        owned = np.flatnonzero(owner == r)
        tgt = []
        for s in range(link_type.shape[0]):
            has = link_type[s, owned] != 0
            tgt.append(link_index[s, owned][has].astype(np.int64))
        tgt = np.unique(np.concatenate(tgt)) if tgt else np.zeros(0, np.int64)
        ghosts = tgt[owner[tgt] != r]
        l2g = np.concatenate([owned[surface_flag[owned]], ghosts[surface_flag[ghosts]],
                              owned[~surface_flag[owned]], ghosts[~surface_flag[ghosts]]])
        out.append(RankGraph(rank=r, owned=owned, ghosts=ghosts, local_to_global=l2g,
                             n_surface_local=int(surface_flag[owned].sum() + surface_flag[ghosts].sum()),
                             peers=[], send=[], recv=[]))
    for r, g in enumerate(out):
        g2l = g.global_to_local()
        for p in sorted(set(owner[g.ghosts].tolist())):
            mine_needed_by_p = np.intersect1d(out[p].ghosts, g.owned)          # ascending global id on both sides
            from_p = g.ghosts[owner[g.ghosts] == p]
            g.peers.append(int(p))
            g.recv.append(np.array([g2l[int(x)] for x in np.sort(from_p)], np.uint32))
            g.send.append(np.array([g2l[int(x)] for x in mine_needed_by_p], np.uint32))
        for p in range(world):                                                  # peers that only receive from me
            if p != r and p not in g.peers:
                need = np.intersect1d(out[p].ghosts, g.owned)
                if need.size:
                    g.peers.append(int(p))
                    g.send.append(np.array([g2l[int(x)] for x in need], np.uint32))
                    g.recv.append(np.zeros(0, np.uint32))
    return out
