"""The Redis command surface of the reference (src/lib.rs, src/types.rs; Readme.md:49-189; cmd.sh) served by
libredis_hnsw_b200.so inside the fake module host, with the index on the GPU.  No redis-server exists in this image, so
the host is tests/fake_redis/fake_redis_host.cpp (same function-table ABI)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
import redis_host as R  # noqa: E402
from redis_hnsw_b200 import data  # noqa: E402


def _vec(v):
    return " ".join(R.fmt(x) for x in v)


def test_cmd_sh_session():
    """cmd.sh:4-25 — NEW (DIM 128 M 5), 100x NODE.ADD, GET, NODE.GET, SEARCH, 100x NODE.DEL, DEL."""
    cmds = ["HNSW.NEW test1 DIM 128 M 5"]
    for i in range(1, 101):
        cmds.append("HNSW.NODE.ADD test1 node%d DATA 128 %s" % (i, " ".join([str(i)] * 128)))
    cmds += ["HNSW.GET test1", "HNSW.NODE.GET test1 node1", "HNSW.SEARCH test1 QUERY 128 " + " ".join(["2"] * 128), "#KEYS"]
    for i in range(1, 101):
        cmds.append("HNSW.NODE.DEL test1 node%d" % i)
    cmds += ["HNSW.GET test1", "HNSW.DEL test1", "#KEYS", "HNSW.GET test1"]
    r = R.run(cmds)
    assert r[0] == {"status": "OK"}                                              # lib.rs:170
    assert all(x == {"status": "OK"} for x in r[1:101])                          # lib.rs:367
    g = R.pairs(r[101])                                                          # types.rs:122-155
    assert list(g) == ["name", "metric", "data_dim", "m", "ef_construction", "level_mult", "node_count", "max_layer",
                       "enterpoint"]
    assert g["name"] == "hnsw.test1" and g["metric"] == "Euclidean" and g["data_dim"] == 128 and g["m"] == 5
    assert g["ef_construction"] == 200 and g["node_count"] == 100               # EFCON default 200 (lib.rs:53)
    assert abs(g["level_mult"] - 1.0 / np.log(5.0)) < 1e-12                      # core.rs:338
    assert g["enterpoint"].startswith("hnsw.test1.node")
    nd = R.pairs(r[102])                                                         # types.rs:322-352
    assert nd["data"] == [1.0] * 128
    assert len(nd["neighbors"]) >= 1 and all(n.startswith("hnsw.test1.node") for n in nd["neighbors"][0])
    s = r[103]                                                                   # lib.rs:485-492: K defaults to 5
    assert s[0] == 5 and len(s) == 6
    assert s[1] == ["similarity", -0.0, "name", "node2"]                         # name = last '.'-segment (core.rs:885-887)
    assert sorted(x[1] for x in s[2:4]) == [-128.0, -128.0] and sorted(x[3] for x in s[2:4]) == ["node1", "node3"]
    assert len(r[104]) == 101 and ["hnsw.test1", "hnswindex"] in r[104] and ["hnsw.test1.node7", "hnswnodet"] in r[104]
    assert all(x == 1 for x in r[105:205])                                       # lib.rs:406: integer 1
    g2 = R.pairs(r[205])
    assert g2["node_count"] == 0 and g2["enterpoint"] is None
    assert r[206] == 1                                                           # lib.rs:226
    assert r[207] == []
    assert r[208] == {"error": "Index: hnsw.test1 does not exist"}


def test_reference_core_kat_and_error_texts():
    """core_tests.rs:7-53 through the commands; error strings carry the reference's renderings."""
    cmds = ["HNSW.NEW foo DIM 4 M 5 EFCON 16"]
    for i in range(100):
        cmds.append("HNSW.NODE.ADD foo node%d DATA 4 %d %d %d %d" % (i, i, i, i, i))
    cmds += ["HNSW.SEARCH foo K 5 QUERY 4 10 10 10 10",
             "HNSW.NEW foo DIM 4",
             "HNSW.NODE.ADD foo node7 DATA 4 0 0 0 0",
             "HNSW.NODE.ADD foo x DATA 3 0 0 0",
             "HNSW.SEARCH foo QUERY 5 0 0 0 0 0",
             "HNSW.NODE.DEL foo nope",
             "HNSW.SEARCH foo K 3 EF 64 QUERY 4 50.2 50.2 50.2 50.2",
             "HNSW.SEARCH foo K 0 QUERY 4 1 1 1 1"]
    r = R.run(cmds)
    s = r[101]
    assert s[0] == 5 and [x[1] for x in s[1:]] == [-0.0, -4.0, -4.0, -16.0, -16.0] and s[1][3] == "node10"
    assert r[102] == {"error": "Index: hnsw.foo already exists"}                                        # lib.rs:146-149
    assert r[103] == {"error": 'String("Node: \\"hnsw.foo.node7\\" already exists")'}                   # core.rs:408 + :41-45
    assert r[104] == {"error": 'String("data dimension: 3 does not match Index")'}                      # core.rs:390
    assert r[105] == {"error": 'String("data dimension: 5 does not match Index")'}                      # core.rs:479
    assert r[106] == {"error": 'String("Node: \\"hnsw.foo.nope\\" does not exist")'}                    # core.rs:421
    assert [x[3] for x in r[107][1:]] == ["node50", "node51", "node49"]
    assert r[108] == [0]


def _build_cmds(idx, x, m, efc):
    cmds = ["HNSW.NEW %s DIM %d M %d EFCON %d" % (idx, x.shape[1], m, efc)]
    for i in range(x.shape[0]):
        cmds.append("HNSW.NODE.ADD %s n%d DATA %d %s" % (idx, i, x.shape[1], _vec(x[i])))
    return cmds


def _graph_from_replies(index_reply, node_replies, names, dim):
    """IndexRedis + NodeRedis replies -> the flat graph the oracle imports (ids = position in `names`)."""
    g = R.pairs(index_reply)
    ids = {n: i for i, n in enumerate(names)}
    levels, offs, nbrs, vecs = [], [0], [], []
    for rep in node_replies:
        nd = R.pairs(rep)
        vecs.append(np.asarray(nd["data"], np.float32))
        levels.append(len(nd["neighbors"]) - 1)
        for layer in nd["neighbors"]:
            nbrs += [ids[n] for n in layer]
            offs.append(len(nbrs))
    return np.stack(vecs), dict(levels=np.asarray(levels, np.int32), row_offs=np.asarray(offs, np.uint64),
                                nbrs=np.asarray(nbrs, np.uint32), entry=ids[g["enterpoint"]], max_layer=g["max_layer"])


def test_search_replies_equal_the_oracle_and_survive_an_rdb_round_trip(tmp_path):
    """Build through HNSW.NODE.ADD, read the graph back through HNSW.GET / HNSW.NODE.GET, and check every HNSW.SEARCH
    reply against the CPU oracle searching that same graph (names and f64-widened sims exact); then save the two data
    types, boot a fresh host, load them (types.rs:180-241, 377-408; lazy rebuild lib.rs:229-315) and get the same
    replies; then keep mutating."""
    n, dim, m, efc, nq = 500, 32, 5, 64, 40
    x, q = data.uniform(n, dim, seed=5, n_queries=nq)
    names = ["hnsw.idx.n%d" % i for i in range(n)]
    rdb = str(tmp_path / "dump.fake_rdb")
    search = ["HNSW.SEARCH idx K 10 QUERY %d %s" % (dim, _vec(v)) for v in q]
    msearch = "HNSW.MSEARCH idx K 10 QUERIES %d %d %s" % (nq, dim, " ".join(_vec(v) for v in q))
    cmds = _build_cmds("idx", x, m, efc) + ["HNSW.GET idx"] + ["HNSW.NODE.GET idx n%d" % i for i in range(n)] + search + \
        [msearch, "#SAVE " + rdb]
    r = R.run(cmds)
    assert all(v == {"status": "OK"} for v in r[:n + 1])
    index_reply, node_replies = r[n + 1], r[n + 2:2 * n + 2]
    replies = r[2 * n + 2:2 * n + 2 + nq]
    assert r[2 * n + 2 + nq] == replies                                          # MSEARCH == nq SEARCH replies
    assert r[-1] == n + 1                                                        # keys written
    vecs, g = _graph_from_replies(index_reply, node_replies, names, dim)
    assert np.array_equal(vecs, x)
    orc = oracle.Oracle(dim, m, efc)
    orc.import_graph(vecs, g)
    oids, osims, ocnt, ost, _ = orc.search_batch(q, 10)                          # ef = ef_construction (core.rs:485)
    checked = 0
    for i in range(nq):
        if ost[i, 3]:
            continue                                                             # tie between different nodes
        want = [int(ocnt[i])] + [["similarity", float(osims[i, j]), "name", "n%d" % oids[i, j]] for j in range(int(ocnt[i]))]
        assert replies[i] == want
        checked += 1
    assert checked > nq * 0.9

    # fresh process: load the RDB, same answers, graph intact, mutations continue
    extra, _ = data.uniform(20, dim, seed=77, n_queries=1)
    cmds2 = ["#LOAD " + rdb, "#KEYS", "HNSW.NODE.GET idx n3"] + search + ["HNSW.GET idx", "HNSW.NODE.GET idx n3"]
    cmds2 += ["HNSW.NODE.ADD idx e%d DATA %d %s" % (i, dim, _vec(extra[i])) for i in range(20)]
    cmds2 += ["HNSW.NODE.DEL idx n10", "HNSW.GET idx", "HNSW.SEARCH idx K 1 QUERY %d %s" % (dim, _vec(extra[4])),
              "HNSW.DEL idx", "#KEYS"]
    r2 = R.run(cmds2)
    assert r2[0] == n + 1 and len(r2[1]) == n + 1
    assert r2[2] == node_replies[3]                                              # served from the loaded record (no rebuild yet)
    assert r2[3:3 + nq] == replies                                               # lazy rebuild, identical search replies
    assert R.pairs(r2[3 + nq]) == R.pairs(index_reply)
    assert r2[4 + nq] == node_replies[3]                                         # now served from the device
    base = 5 + nq
    assert all(v == {"status": "OK"} for v in r2[base:base + 20])
    assert r2[base + 20] == 1
    g2 = R.pairs(r2[base + 21])
    assert g2["node_count"] == n + 19
    assert r2[base + 22][1][1] == -0.0 and r2[base + 22][1][3] == "e4"
    assert r2[base + 23] == 1 and r2[base + 24] == []
    os.remove(rdb)


@pytest.mark.parametrize("events", [True, False])
def test_bgsave_in_a_forked_child_sees_current_records(tmp_path, events):
    """Redis writes RDB files from a fork()ed child, where the parent's CUDA context is unusable.  With server events the
    module snapshots every record to host memory when the persistence event fires (before the fork); without them it keeps
    the records current after every mutation, as the reference does (lib.rs:351-365).  Either way the child must write the
    CURRENT graph: mutate after a first save, BGSAVE, load in a fresh process, compare against the parent's replies."""
    n, dim, m, efc = 300, 32, 5, 48
    x, q = data.uniform(n + 30, dim, seed=8, n_queries=10)
    env = None if events else {"FAKE_REDIS_NO_EVENTS": "1"}
    rdb1, rdb2 = str(tmp_path / "a.fake_rdb"), str(tmp_path / "b.fake_rdb")
    probe = ["HNSW.GET idx"] + ["HNSW.NODE.GET idx n%d" % i for i in (0, 5, 17, n + 3, n + 29)] + \
        ["HNSW.SEARCH idx K 10 QUERY %d %s" % (dim, _vec(v)) for v in q]
    cmds = _build_cmds("idx", x[:n], m, efc) + ["#BGSAVE " + rdb1]
    cmds += ["HNSW.NODE.ADD idx n%d DATA %d %s" % (i, dim, _vec(x[i])) for i in range(n, n + 30)]
    cmds += ["HNSW.NODE.DEL idx n%d" % i for i in (7, 8, 9)]
    cmds += ["#BGSAVE " + rdb2] + probe
    r = R.run(cmds, env=env)
    assert r[n + 1] == {"status": "Background saving done"}
    assert r[n + 35] == {"status": "Background saving done"}
    want = r[n + 36:]
    got = R.run(["#LOAD " + rdb2] + probe)
    assert got[0] == n + 30 - 3 + 1
    assert got[1:] == want
    first = R.run(["#LOAD " + rdb1, "HNSW.GET idx"])
    assert first[0] == n + 1 and R.pairs(first[1])["node_count"] == n


@pytest.mark.parametrize("fast,events", [(1, True), (0, True), (1, False)])
def test_node_madd_bulk_load(tmp_path, fast, events):
    """HNSW.NODE.MADD (extension): a NODE.ADD stream in one command.  FAST 0 must give the graph of the one-by-one stream
    when the levels are the same — the engine draws them from the same generator in both cases — and every form must
    leave a consistent keyspace that survives BGSAVE + reload."""
    n, dim, m, efc = 400, 32, 5, 48
    x, q = data.uniform(n, dim, seed=21, n_queries=8)
    # the engine seeds its level RNG from entropy (core.rs:344); both processes of the FAST 0 comparison pin it instead
    env = dict({"HNSW_LEVEL_SEED": "7"}, **({} if events else {"FAKE_REDIS_NO_EVENTS": "1"}))
    rdb = str(tmp_path / "m.fake_rdb")
    madd = "HNSW.NODE.MADD idx FAST %d NODES %d %s DATA %d %d %s" % (
        fast, n - 1, " ".join("n%d" % i for i in range(1, n)), n - 1, dim, " ".join(_vec(v) for v in x[1:]))
    probe = ["HNSW.GET idx"] + ["HNSW.NODE.GET idx n%d" % i for i in (0, 1, 77, n - 1)] + \
        ["HNSW.SEARCH idx K 5 EF 48 QUERY %d %s" % (dim, _vec(v)) for v in q]
    cmds = ["HNSW.NEW idx DIM %d M %d EFCON %d" % (dim, m, efc), "HNSW.NODE.ADD idx n0 DATA %d %s" % (dim, _vec(x[0])), madd,
            "HNSW.NODE.MADD idx NODES 1 n5 DATA 1 %d %s" % (dim, _vec(x[5])), "#KEYS", "#BGSAVE " + rdb] + probe
    r = R.run(cmds, env=env)
    assert r[2] == n - 1
    assert r[3] == {"error": 'String("Node: \\"hnsw.idx.n5\\" already exists")'}
    assert len(r[4]) == n + 1
    assert R.pairs(r[6])["node_count"] == n
    got = R.run(["#LOAD " + rdb] + probe)
    assert got[0] == n + 1 and got[1:] == r[6:]
    if fast == 0:                                                  # same levels, same order -> same graph as one-by-one
        one = R.run(_build_cmds("idx", x, m, efc) + probe, env=env)
        assert one[n + 1:] == r[6:]


def test_hand_derived_reference_rdb_loads_into_the_engine():
    """The reference-format payloads of tests/golden/rdb_5node.fake_rdb (derived by hand from types.rs:243-284, 410-428, see
    tests/golden/make_rdb_fixture.py) rebuild a working device index: the lazy rebuild of lib.rs:229-315 from records this
    repo did not write.  core_tests.rs:45-53 in miniature: the nearest neighbours of [2;4] are n2 (sim -0), then n1 / n3
    (sim -4 each)."""
    import os

    fixture = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rdb_5node.fake_rdb")
    r = R.run(["#LOAD " + fixture, "HNSW.GET kat5", "HNSW.SEARCH kat5 K 3 QUERY 4 2 2 2 2", "HNSW.GET empty",
               "HNSW.SEARCH empty K 3 QUERY 8 0 0 0 0 0 0 0 0", "HNSW.NODE.ADD kat5 n5 DATA 4 5 5 5 5", "HNSW.GET kat5",
               "HNSW.NODE.GET kat5 n5", "HNSW.NODE.ADD empty first DATA 8 1 1 1 1 1 1 1 1", "HNSW.GET empty"])
    assert r[0] == 7
    g = R.pairs(r[1])
    assert (g["name"], g["metric"], g["data_dim"], g["m"], g["ef_construction"], g["node_count"], g["max_layer"], g["enterpoint"]) == \
        ("hnsw.kat5", "Euclidean", 4, 5, 16, 5, 1, "hnsw.kat5.n2")
    assert r[2][0] == 3
    hits = [R.pairs(x) for x in r[2][1:]]
    assert [h["similarity"] for h in hits] == [-0.0, -4.0, -4.0] and hits[0]["name"] == "n2"
    assert {hits[1]["name"], hits[2]["name"]} == {"n1", "n3"}
    e = R.pairs(r[3])
    assert (e["node_count"], e["max_layer"], e["enterpoint"]) == (0, 0, None)      # "null" (types.rs:234-237, 277-283)
    assert r[4] == [0]                                                               # core.rs:481-483
    assert r[5] == {"status": "OK"} and R.pairs(r[6])["node_count"] == 6
    rec = R.pairs(r[7])
    assert rec["data"] == [5.0] * 4 and set(rec["neighbors"][0]) == {"hnsw.kat5.n%d" % i for i in range(5)}
    assert r[8] == {"status": "OK"}
    e2 = R.pairs(r[9])
    assert (e2["node_count"], e2["enterpoint"]) == (1, "hnsw.empty.first")
