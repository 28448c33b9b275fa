"""Host-side helpers for the sharded-map mode (BASELINE configs[4]).

The local map is split into x-slabs, one per rank, each extended by a halo of 1 m: the reference only accepts a
5-NN correspondence when all five neighbours lie within 1 m of the query (LM:762, 869), so a rank that owns a query
(by the query's map-frame x) finds every neighbour that can matter inside its slab + halo.  Each rank evaluates its
own queries and contributes 28 doubles; one all-reduce gives every rank the full normal equations.
"""
import numpy as np

HALO = 1.0  # metres; sqrt of the 1.0 m^2 acceptance gate


def slab_edges(x_min, x_max, world):
    """world+1 ascending edges; the outer two are infinite so nothing falls outside."""
    e = np.linspace(x_min, x_max, world + 1).astype(np.float64)
    e[0], e[-1] = -np.inf, np.inf
    return e


def shard_map(cloud4, edges, rank, halo=HALO):
    """Points of `cloud4` whose x lies in the rank's slab extended by the halo (order preserved)."""
    x = cloud4[:, 0].astype(np.float64)
    return cloud4[(x >= edges[rank] - halo) & (x < edges[rank + 1] + halo)]


def associate_to_map(stack4, T):
    """pointAssociateToMap (LM:244-262) in fp32 numpy — used only to ROUTE queries to their owner."""
    T = np.asarray(T, np.float32)
    srx, crx, sry, cry, srz, crz = (np.float32(f(T[i])) for i in (0, 1, 2) for f in (np.sin, np.cos))
    x, y, z = stack4[:, 0], stack4[:, 1], stack4[:, 2]
    x1 = crz * x - srz * y
    y1 = srz * x + crz * y
    y2 = crx * y1 - srx * z
    z2 = srx * y1 + crx * z
    out = stack4.copy()
    out[:, 0] = cry * x1 + sry * z2 + T[3]
    out[:, 1] = y2 + T[4]
    out[:, 2] = -sry * x1 + cry * z2 + T[5]
    return out


def route_queries(stack4, T, edges, rank):
    """Stack points (sensor frame) owned by `rank`: those whose map-frame x falls into the rank's slab."""
    xm = associate_to_map(stack4, T)[:, 0].astype(np.float64)
    return stack4[(xm >= edges[rank]) & (xm < edges[rank + 1])]
