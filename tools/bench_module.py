#!/usr/bin/env python
"""Per-command latency of the Redis module inside the fake host: the host reports the wall time of every command
handler (FAKE_REDIS_TIMING=1), i.e. argument parsing inside the module + engine call + reply building, without the
script parsing of the test harness.  100K x 128 index loaded with HNSW.NODE.MADD (FAST)."""
import collections, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import redis_host as R
from redis_hnsw_b200 import data

n, dim, nq = 100_000, 128, 1000
x, q = data.lowrank(n + 500, dim, seed=123, n_queries=nq)
vec = lambda v: " ".join(R.fmt(t) for t in v)
cmds = ["HNSW.NEW idx DIM %d M 16 EFCON 200" % dim]
for s in range(0, n, 10000):
    cmds.append("HNSW.NODE.MADD idx NODES %d %s DATA %d %d %s" % (10000, " ".join("n%d" % i for i in range(s, s + 10000)), 10000, dim,
                                                                   " ".join(vec(v) for v in x[s:s + 10000])))
cmds += ["HNSW.SEARCH idx K 10 EF 64 QUERY %d %s" % (dim, vec(v)) for v in q]
cmds += ["HNSW.MSEARCH idx K 10 EF 64 QUERIES %d %d %s" % (nq, dim, " ".join(vec(v) for v in q))]
cmds += ["HNSW.NODE.ADD idx e%d DATA %d %s" % (i, dim, vec(x[n + i])) for i in range(500)]
cmds += ["HNSW.NODE.DEL idx n%d" % i for i in range(0, 3000, 10)]
cmds += ["HNSW.NODE.GET idx n%d" % i for i in range(5000, 5200)]
R.build()
p = subprocess.run([R.HOST, R.MODULE], input="\n".join(cmds) + "\n", capture_output=True, text=True,
                   env=dict(os.environ, FAKE_REDIS_TIMING="1"), timeout=3000)
assert p.returncode == 0, p.stderr[-2000:]
t = collections.defaultdict(list)
for line in p.stderr.splitlines():
    if line.startswith("T "):
        _, name, us = line.split()
        t[name].append(float(us))
out = {}
for name, v in t.items():
    v = np.asarray(v)
    out[name] = {"n": int(v.size), "p50_us": float(np.percentile(v, 50)), "p99_us": float(np.percentile(v, 99))}
out["hnsw.msearch"]["us_per_query"] = out["hnsw.msearch"]["p50_us"] / nq
out["hnsw.node.madd"]["nodes_per_s"] = 10000 / (out["hnsw.node.madd"]["p50_us"] / 1e6)
import json
print(json.dumps(out))
