#!/usr/bin/env python
"""Golden fingerprint of the ORACLE-built graph for 100 000 x 128-d, M=16, ef_construction=200 (BASELINE configs[1]
parameters at a size the CPU oracle builds in minutes): the device builders claim list-for-list identity with the
reference's sequential NODE.ADD stream (core.rs:489-599), and at this size the oracle is too slow to be rebuilt inside a
GPU test, so its graph is pinned here once.

    python tests/golden/make_graph_fingerprint.py        # ~6 min on one core; writes graph_fingerprint_100k.json

The fingerprint holds, for every prefix checkpoint: sha256 of the exported arrays (levels, row offsets, neighbour ids,
enterpoint, max_layer) in list order, the same with every list taken as a set, order-sensitive digests of every block of
1000 rows (a mismatch is localised), the oracle's tie counters, and a few literal adjacency lists.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

N, DIM, M, EFC = 100_000, 128, 16, 200
CHECKPOINTS = (10_000, 30_000, 60_000, 100_000)
# The data seed was chosen (first of 123, 124, ... 127 tried) so that the oracle meets NO tie across a select_neighbors cut
# during the build: where two different nodes tie for the m-th place the reference's choice is an accident of BinaryHeap
# layout, and a fixture must not pin accidents.  Seeds 123 / 124 / 125 meet 1 / 2 / 4 such ties,
# 126 and 127 none (the `select_ties` field of every checkpoint).
DATA_SEED = int(os.environ.get("FP_DATA_SEED", "126"))
LEVEL_SEED = 42


def _rows(g):
    offs = np.ascontiguousarray(g["row_offs"], np.uint64)
    nbrs = np.ascontiguousarray(g["nbrs"], np.uint32)
    return offs, nbrs


def graph_digest(g, as_sets=False):
    """sha256 of the exported graph.  as_sets = True hashes every adjacency list as a SET (ids ascending): two graphs that
    differ only in the order of neighbours whose sims tie exactly (the reference leaves that order to
    BinaryHeap internals) have the same set digest."""
    offs, nbrs = _rows(g)
    if as_sets:
        nbrs = nbrs.copy()
        row_of = np.repeat(np.arange(offs.size - 1, dtype=np.int64), np.diff(offs).astype(np.int64))
        order = np.lexsort((nbrs, row_of))
        nbrs = nbrs[order]
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(g["levels"], np.int32).tobytes())
    h.update(offs.tobytes())
    h.update(nbrs.tobytes())
    h.update(np.array([g["entry"], g["max_layer"]], np.int64).tobytes())
    return h.hexdigest()


def block_digests(g, block=1000):
    """Order-sensitive digest of every block of `block` adjacency rows (16 hex digits each): localises a difference."""
    offs, nbrs = _rows(g)
    out = []
    for r0 in range(0, offs.size - 1, block):
        r1 = min(offs.size - 1, r0 + block)
        h = hashlib.sha256()
        h.update((offs[r0:r1 + 1] - offs[r0]).tobytes())
        h.update(nbrs[int(offs[r0]):int(offs[r1])].tobytes())
        out.append(h.hexdigest()[:16])
    return out


def dataset():
    from redis_hnsw_b200 import data

    x, _ = data.lowrank(N, DIM, r=16, seed=DATA_SEED)
    levels = data.draw_levels(N, M, seed=LEVEL_SEED)
    return x, levels


def main():
    import oracle

    x, levels = dataset()
    orc = oracle.Oracle(DIM, M, EFC)
    out = {"n": N, "dim": DIM, "m": M, "ef_construction": EFC, "dataset": "lowrank r=16 sigma=0.05 seed=%d" % DATA_SEED,
           "levels": "draw_levels seed=%d" % LEVEL_SEED, "generator": "oracle/hnsw_oracle.cpp via tests/golden/make_graph_fingerprint.py",
           "checkpoints": {}}
    done = 0
    t0 = time.time()
    ties = 0
    oracle.cut_ties(reset=True)
    for cp in CHECKPOINTS:
        st = orc.add_batch(x[done:cp], levels[done:cp])
        ties += int(st[3])
        done = cp
        g = orc.export_graph()
        deg0 = np.diff(g["row_offs"])
        out["checkpoints"][str(cp)] = {"sha256": graph_digest(g), "set_sha256": graph_digest(g, as_sets=True),
                                       "row_blocks": block_digests(g), "order_ties": oracle.cut_ties()[2],
                                       "edges": int(g["nbrs"].size), "rows": int(deg0.size),
                                       "max_degree": int(deg0.max()), "entry": g["entry"], "max_layer": g["max_layer"],
                                       # equal sims of different nodes across a cut so far: (select_neighbors, w evictions)
                                       "select_ties": oracle.cut_ties()[0], "evict_ties": oracle.cut_ties()[1]}
        print(cp, out["checkpoints"][str(cp)], "%.0f s" % (time.time() - t0), flush=True)
    out["heap_ties_met_by_the_oracle"] = ties
    out["sample_lists"] = {str(i): [int(v) for v in orc.node_neighbors(i, 0)] for i in (0, 1, 777, 50_000, 99_999)}
    name = os.environ.get("FP_OUT", "graph_fingerprint_100k.json")
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), name), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
