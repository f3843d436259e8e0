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
