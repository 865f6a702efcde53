"""Multi-GPU sharding of the render loop (one process per GPU).

Every pixel sample is independent given (scene, camera, params, seed, y, x, epoch), so the path shards
without any data-path exchange:
  * epochs   (stochastic pass, main.rs:1129): rank g renders epochs [g*E/G, (g+1)*E/G) of the full frame into
             its own {sum.rgb, count} buffer; ONE sum all-reduce of the buffers ends the render.
  * rows     (Whitted pass, main.rs:1090): rank g renders rows [g*H/G, (g+1)*H/G); the disjoint bands are
             gathered (a sum all-reduce of zero-initialised frames is a gather).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: returns (begin, count) of [rank*total/world, (rank+1)*total/world)."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    b = (rank * total) // world
    e = ((rank + 1) * total) // world
    return b, e - b


def all_shards(total: int, world: int):
    return [shard_range(total, r, world) for r in range(world)]
