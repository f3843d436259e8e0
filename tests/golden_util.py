"""Loader of the committed golden vectors (tests/golden/)."""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPHS = ("d32_m5", "d128_m16", "d20_m6")
METRIC_DIMS = (32, 128, 768, 20, 33)


def kats():
    with open(os.path.join(HERE, "reference_kats.json")) as f:
        return json.load(f)


def load():
    return np.load(os.path.join(HERE, "hnsw_golden.npz"))


def graph(z, name, prefix=""):
    """The flat graph dict of DeviceIndex.load_graph / Oracle.import_graph."""
    e = z[name + "_" + prefix + "entry"]
    return dict(levels=z[name + "_" + prefix + "levels"].astype(np.int32),
                row_offs=z[name + "_" + prefix + "row_offs"].astype(np.uint64),
                nbrs=z[name + "_" + prefix + "nbrs"].astype(np.uint32), entry=int(e[0]), max_layer=int(e[1]))


def same_graph(a, b):
    return (np.array_equal(a["levels"], b["levels"]) and np.array_equal(a["row_offs"], b["row_offs"])
            and np.array_equal(a["nbrs"], b["nbrs"]) and a["entry"] == b["entry"] and a["max_layer"] == b["max_layer"])
