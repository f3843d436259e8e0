#!/usr/bin/env python
"""Where does a SPEC build leave the oracle's graph?  Builds the fingerprint dataset with SPEC and with the one-warp EXACT
kernel side by side in steps, compares the two device graphs at every step, and saves both at the first step where either
differs from the other or from the committed oracle fingerprint (gpurun_out/diag_spec_<nodes>.npz)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_graph_fingerprint as fp  # noqa: E402


def main():
    import redis_hnsw_b200 as r

    lo, hi, step = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "graph_fingerprint_100k.json")))["checkpoints"]
    x, levels = fp.dataset()
    a = r.DeviceIndex(fp.DIM, fp.M, fp.EFC)
    b = r.DeviceIndex(fp.DIM, fp.M, fp.EFC)
    a.reserve(hi)
    b.reserve(hi)
    a.add_batch(x[:lo], levels[:lo], mode=r.BUILD_SPEC)
    ga = a.export_graph()
    b.load_graph(x[:lo], ga)                      # the EXACT stream continues from the SPEC graph (equal to the oracle's at `lo`)
    print("start", lo, fp.graph_digest(ga), gold.get(str(lo), {}).get("sha256"), flush=True)
    done = lo
    while done < hi:
        k = min(step, hi - done)
        a.add_batch(x[done:done + k], levels[done:done + k], mode=r.BUILD_SPEC)
        b.add_batch(x[done:done + k], levels[done:done + k], mode=r.BUILD_EXACT)
        done += k
        ga, gb = a.export_graph(), b.export_graph()
        da, db = fp.graph_digest(ga), fp.graph_digest(gb)
        print(done, "spec", da[:16], "exact", db[:16], "gold", (gold.get(str(done), {}).get("sha256") or "-")[:16], a.build_stats(), flush=True)
        if da != db:
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_spec_%d.npz" % done),
                                **{"spec_" + k2: v for k2, v in ga.items() if isinstance(v, np.ndarray)},
                                **{"exact_" + k2: v for k2, v in gb.items() if isinstance(v, np.ndarray)})
            print("SPEC and EXACT differ at", done, flush=True)
            return
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_spec_%d.npz" % done),
                        **{"spec_" + k2: v for k2, v in ga.items() if isinstance(v, np.ndarray)})


if __name__ == "__main__":
    main()
