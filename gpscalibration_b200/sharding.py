"""Host-side helpers for the sharded-map mode (BASELINE configs[4]).

The local map is split into x-slabs, one per rank, each extended by a halo of 1 m: the reference only accepts a
5-NN correspondence when all five neighbours lie within 1 m of the query (LM:762, 869), so the rank that owns a query
finds every neighbour that can matter inside its slab + halo.  Only the MAP is distributed on the host (once, when it is
loaded).  Which rank evaluates a stack point is decided on the device, every Gauss-Newton iteration, from the point's
map-frame x under the current pose (`loam_shard_set_slab`: every rank holds the whole stack); each rank contributes 28
doubles and the ranks exchange them inside the kernel (`loam_map_optimize`, `loam_map_iter_allreduce`) or through NCCL
(`loam_map_iter_partial`).
"""
import numpy as np

HALO = 1.0  # metres; sqrt of the 1.0 m^2 acceptance gate


def slab_edges(x_min, x_max, world):
    """world+1 ascending float32 edges; the outer two are infinite so nothing falls outside."""
    e = np.linspace(x_min, x_max, world + 1).astype(np.float32)
    e[0], e[-1] = -np.inf, np.inf
    return e


def shard_indices(cloud4, edges, rank, halo=HALO):
    """Indices of the points of `cloud4` whose x lies in the rank's slab extended by the halo (ascending)."""
    x = cloud4[:, 0].astype(np.float64)
    return np.nonzero((x >= float(edges[rank]) - halo) & (x < float(edges[rank + 1]) + halo))[0]


def shard_map(cloud4, edges, rank, halo=HALO):
    """Points of `cloud4` whose x lies in the rank's slab extended by the halo (order preserved)."""
    return cloud4[shard_indices(cloud4, edges, rank, halo)]
