"""CPU restatement of the on-device owner rule of the sharded-map mode (loam_shard_set_slab): a stack point belongs to
the rank whose slab [edges[r], edges[r+1]) contains its map-frame x under pose T (pointAssociateToMap, LM:244-262,
evaluated by the oracle with the reference's fp32 arithmetic)."""
import numpy as np


def owner_mask(orc, stack4, T, edges, rank):
    xm = orc.associate_to_map(stack4, T)[:, 0]
    return (xm >= np.float32(edges[rank])) & (xm < np.float32(edges[rank + 1]))
