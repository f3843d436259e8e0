"""The Redis module on a machine without a GPU: it loads into the (fake) module host, registers exactly the
reference's command table and data types (src/lib.rs:498-514, src/types.rs:157,354), parses arguments, and FAILS LOUDLY
(an error reply, not a crash, not a CPU fallback) when a command needs the device."""
import redis_host as R


def test_onload_registers_the_reference_command_table():
    name, ver, cmds, types = R.run(["#INFO"])[0]
    assert (name, ver) == ("hnsw", 1)                                           # lib.rs:499-500
    table = {c[0]: tuple(c[1:]) for c in cmds}
    ref = {"hnsw.new": "write", "hnsw.get": "readonly", "hnsw.del": "write", "hnsw.search": "readonly",
           "hnsw.node.add": "write", "hnsw.node.get": "readonly", "hnsw.node.del": "write"}   # lib.rs:505-513
    for cmd, flags in ref.items():
        assert table[cmd] == (flags, 0, 0, 0), cmd
    assert set(table) - set(ref) == {"hnsw.msearch", "hnsw.node.madd"}          # the labelled extensions
    assert sorted(types) == [["hnswindex", 0], ["hnswnodet", 0]]                # types.rs:13-14,157,354


def test_argument_errors_and_missing_keys():
    r = R.run(["HNSW.NEW", "HNSW.NEW foo", "HNSW.NEW foo DIM x", "HNSW.NEW foo DIM 4 BOGUS 1", "HNSW.GET foo",
               "HNSW.DEL foo", "HNSW.NODE.GET foo bar", "HNSW.NODE.ADD foo bar DATA 4 1 2 3", "HNSW.NODE.ADD foo bar",
               "HNSW.SEARCH foo K 3 QUERY 2 1 2", "hnsw.node.del foo bar", "HNSW.MSEARCH foo QUERIES 2 2 1 2 3",
               "#KEYS"])
    assert all(R.is_error(x) for x in r[:-1])
    assert r[4]["error"] == "Index: hnsw.foo does not exist"                   # lib.rs:241
    assert r[6]["error"] == "Node: hnsw.foo.bar does not exist"               # lib.rs:440
    assert "more values than were given" in r[7]["error"]
    assert r[9]["error"] == "Index: hnsw.foo does not exist"
    assert r[-1] == []                                                          # nothing was created


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        return
    r = R.run(["HNSW.NEW foo DIM 4 M 5 EFCON 16", "#KEYS"])
    assert R.is_error(r[0]) and "CUDA" in r[0]["error"]
    assert r[1] == []


def test_reference_format_rdb_payloads_load_and_round_trip_byte_for_byte(tmp_path):
    """tests/golden/rdb_5node.fake_rdb holds module values derived BY HAND from the reference's rdb_save callbacks
    (types.rs:243-284, 410-428; encoding version 0) in redis-server's own module-value encoding — a 5-node index with its
    five node keys and an empty index whose enterpoint is the literal "null" (tests/golden/make_rdb_fixture.py has the
    derivation).  The module must load them, serve the records, and write the SAME bytes back: its rdb_load / rdb_save
    are checked against the reference's format here, not against each other."""
    import os
    import sys

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    import make_rdb_fixture as fx

    fixture = os.path.join(gold, "rdb_5node.fake_rdb")
    want = fx.container(fx.build())
    assert open(fixture, "rb").read() == want, "the committed fixture is not what the derivation script produces"
    out = str(tmp_path / "again.fake_rdb")
    nodes, data, nbrs = fx.expected()
    r = R.run(["#LOAD " + fixture, "#KEYS"] + ["HNSW.NODE.GET kat5 " + n for n in nodes] + ["#SAVE " + out])
    assert r[0] == 7
    assert r[1] == [["hnsw.empty", "hnswindex"], ["hnsw.kat5", "hnswindex"]] + [["hnsw.kat5." + n, "hnswnodet"] for n in nodes]
    for n, reply in zip(nodes, r[2:7]):                       # served from the loaded record: no device needed
        rec = R.pairs(reply)
        assert rec["data"] == data[n]
        assert rec["neighbors"] == [["hnsw.kat5." + x for x in layer] for layer in nbrs[n]]
    assert r[7] == 7
    assert open(out, "rb").read() == want                     # rdb_save(rdb_load(x)) == x, byte for byte
