#!/usr/bin/env python
"""Single-query latency probe for ncu: builds the headline index and issues a few one-query HNSW.SEARCH calls.

    ncu --set full --import-source on --clock-control none -k regex:search_knn2 --launch-skip 20 --launch-count 1 \
        -o gpurun_out/prof_lat python tools/lat_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import redis_hnsw_b200 as r

    wl = sys.argv[1] if len(sys.argv) > 1 else "1Mx128_M16_efc200"
    n, dim, m, efc, _, _ = bench.WORKLOADS[wl]
    x, q, levels = bench.make_data(wl, 64)
    dev = r.DeviceIndex(dim, m, efc)
    dev.reserve(n)
    dev.add_batch(x, levels, mode=r.BUILD_FAST)
    for i in range(40):
        dev.search(q[i], 10, ef=64)


if __name__ == "__main__":
    main()
