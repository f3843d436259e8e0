#!/usr/bin/env python
"""Byte-exact RDB payloads of a 5-node `hnswindex` key, its five `hnswnodet` keys and an EMPTY index, derived by hand
from the reference's rdb_save callbacks — not produced by this repo's module:

  hnswindex   src/types.rs:243-284   string name, string metric ("Euclidean", types.rs:20-27 / core.rs:332), five u64
              (data_dim, m, m_max, m_max_0, ef_construction), double level_mult, u64 node_count, u64 max_layer,
              u64 n_layers x (u64 n, n strings), u64 n_nodes x string, string enterpoint — the literal "null" when the
              index has none (types.rs:277-283)
  hnswnodet   src/types.rs:410-428   u64 n, n floats, u64 n_layers x (u64 n, n strings)
  both        encoding version 0 (types.rs:13-14)

wrapped the way redis-server writes a module-typed value (rdb.c rdbSaveObject, RDB_TYPE_MODULE_2; module.c
RM_Save{Unsigned,Double,Float,StringBuffer}): module id, then one opcode-prefixed item per Save* call, then the EOF
opcode — see the encoding summary in tests/fake_redis/fake_redis_host.cpp.

The graph itself (who links to whom, in which order) is the reference's for NODE.ADD n0..n4 with data [i; 4], M=5,
EFCON=16 and level draws 0,0,1,0,0, as restated by oracle/ (5 nodes, everyone links to everyone already present, lists
in insertion / nearest-first order: see `expected()`); every byte of the framing comes from the rules above.

    python tests/golden/make_rdb_fixture.py      # rewrites tests/golden/rdb_5node.fake_rdb and prints the hex of each value

The file framing (count, then key / payload pairs with 8-byte little-endian lengths) is the fake host's #SAVE / #LOAD
container; the payloads are what a real dump.rdb holds after the key name.
"""
import math
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
CHARS = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_"


def rdb_len(v):                                   # rdbSaveLen
    if v < 1 << 6:
        return bytes([v])
    if v < 1 << 14:
        return bytes([0x40 | (v >> 8), v & 0xFF])
    if v <= 0xFFFFFFFF:
        return b"\x80" + struct.pack(">I", v)
    return b"\x81" + struct.pack(">Q", v)


def module_id(name, encver):                      # moduleTypeEncodeId
    i = 0
    for ch in name:
        i = (i << 6) | CHARS.index(ch)
    return (i << 10) | encver


def u(v):                                         # RedisModule_SaveUnsigned: opcode 2
    return rdb_len(2) + rdb_len(v)


def d(v):                                         # RedisModule_SaveDouble: opcode 4, 8 bytes little-endian
    return rdb_len(4) + struct.pack("<d", v)


def f(v):                                         # RedisModule_SaveFloat: opcode 3, 4 bytes little-endian
    return rdb_len(3) + struct.pack("<f", v)


def s(text):                                      # RedisModule_SaveStringBuffer: opcode 5, raw string
    b = text.encode()
    return rdb_len(5) + rdb_len(len(b)) + b


def value(type_name, items):
    return rdb_len(module_id(type_name, 0)) + b"".join(items) + rdb_len(0)


def expected():
    """The 5-node index of the fixture as the reference would hold it (oracle/: NODE.ADD n0..n4, data [i;4], M=5,
    EFCON=16, levels 0,0,1,0,0): lists per node per level, names relative to the index."""
    nodes = ["n0", "n1", "n2", "n3", "n4"]
    data = {n: [float(i)] * 4 for i, n in enumerate(nodes)}
    nbrs = {"n0": [["n1", "n2", "n3", "n4"]],
            "n1": [["n0", "n2", "n3", "n4"]],
            "n2": [["n1", "n0", "n3", "n4"], []],
            "n3": [["n2", "n1", "n0", "n4"]],
            "n4": [["n3", "n2", "n1", "n0"]]}
    return nodes, data, nbrs


def index_value(name, dim, m, efcon, node_names, layers, enterpoint):
    items = [s(name), s("Euclidean"), u(dim), u(m), u(m), u(2 * m), u(efcon), d(1.0 / math.log(float(m))),
             u(len(node_names)), u(max(0, len(layers) - 1)), u(len(layers))]
    for layer in layers:
        items.append(u(len(layer)))
        items += [s(n) for n in layer]
    items.append(u(len(node_names)))
    items += [s(n) for n in node_names]
    items.append(s(enterpoint if enterpoint is not None else "null"))
    return value("hnswindex", items)


def node_value(vec, neighbor_layers):
    items = [u(len(vec))] + [f(v) for v in vec] + [u(len(neighbor_layers))]
    for layer in neighbor_layers:
        items.append(u(len(layer)))
        items += [s(n) for n in layer]
    return value("hnswnodet", items)


def build():
    nodes, data, nbrs = expected()
    full = lambda n: "hnsw.kat5." + n                                   # lib.rs:342-343
    keys = {}
    # layers: a node is listed in the set of its TOP level only (core.rs:596): level 0 -> n0 n1 n3 n4, level 1 -> n2
    keys["hnsw.kat5"] = index_value("hnsw.kat5", 4, 5, 16, [full(n) for n in nodes],
                                    [[full(n) for n in ("n0", "n1", "n3", "n4")], [full("n2")]], full("n2"))
    for n in nodes:
        keys[full(n)] = node_value(data[n], [[full(x) for x in layer] for layer in nbrs[n]])
    # an index that never held a node: no layers (core.rs:341), no nodes, enterpoint "null" (types.rs:277-283)
    keys["hnsw.empty"] = index_value("hnsw.empty", 8, 5, 200, [], [], None)
    return keys


def container(keys):
    out = struct.pack("<Q", len(keys))
    for k in sorted(keys):
        kb = k.encode()
        out += struct.pack("<Q", len(kb)) + kb + struct.pack("<Q", len(keys[k])) + keys[k]
    return out


if __name__ == "__main__":
    keys = build()
    with open(os.path.join(HERE, "rdb_5node.fake_rdb"), "wb") as fh:
        fh.write(container(keys))
    for k in sorted(keys):
        print(k, len(keys[k]), keys[k].hex())
