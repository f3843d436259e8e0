"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/hnsw_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (redis_hnsw_b200/) never does.  See the header of hnsw_oracle.cpp for what is restated
and how it is pinned (reference KATs: src/hnsw/metrics_tests.rs:4-33, src/hnsw/core_tests.rs:7-81).
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only, no reference build system)."""
    src = os.path.join(_HERE, "hnsw_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, u32p, u64p, i32p, i64p = (C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int64))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        for name in ("orc_euclidean", "orc_sim_avx", "orc_sim_avx_portable", "orc_sim_scalar"):
            f = getattr(L, name)
            f.restype = C.c_float
            f.argtypes = [fp, fp, C.c_uint64]
        L.orc_euclidean_batch.argtypes = [fp, fp, C.c_uint64, C.c_uint64, fp]
        L.orc_cut_ties.restype = C.c_uint64
        L.orc_cut_ties.argtypes = []
        L.orc_cut_ties_reset.argtypes = []
        L.orc_evict_ties.restype = C.c_uint64
        L.orc_evict_ties.argtypes = []
        L.orc_order_ties.restype = C.c_uint64
        L.orc_order_ties.argtypes = []
        L.orc_level_from_u.restype = C.c_int
        L.orc_level_from_u.argtypes = [C.c_double, C.c_int]
        L.orc_add.restype = C.c_int64
        L.orc_add.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, u64p]
        L.orc_add_batch.argtypes = [C.c_void_p, C.c_uint64, fp, i32p, u64p]
        L.orc_delete.restype = C.c_int
        L.orc_delete.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_touched.restype = C.c_uint64
        L.orc_touched.argtypes = [C.c_void_p, u32p, C.c_uint64]
        L.orc_search.restype = C.c_int
        L.orc_search.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, u32p, fp, u64p]
        L.orc_search_batch.restype = C.c_double
        L.orc_search_batch.argtypes = [C.c_void_p, C.c_uint64, fp, C.c_int, C.c_int, u32p, fp, u32p, u64p, C.c_int]
        L.orc_search_level.restype = C.c_int
        L.orc_search_level.argtypes = [C.c_void_p, fp, C.c_uint32, C.c_int, C.c_int, u32p, fp, C.c_int]
        L.orc_params.argtypes = [C.c_void_p, i64p]
        L.orc_level_mult.restype = C.c_double
        L.orc_level_mult.argtypes = [C.c_void_p]
        L.orc_n_ids.restype = C.c_uint64
        L.orc_n_ids.argtypes = [C.c_void_p]
        L.orc_node_level.restype = C.c_int
        L.orc_node_level.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_node_n_levels.restype = C.c_int
        L.orc_node_n_levels.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_node_neighbors.restype = C.c_uint64
        L.orc_node_neighbors.argtypes = [C.c_void_p, C.c_uint32, C.c_int, u32p, C.c_uint64]
        L.orc_node_vector.argtypes = [C.c_void_p, C.c_uint32, fp]
        L.orc_graph_sizes.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.orc_export.argtypes = [C.c_void_p, i32p, u64p, u32p, i64p, i32p]
        L.orc_import.argtypes = [C.c_void_p, C.c_uint64, fp, i32p, u64p, u32p, C.c_int64, C.c_int32]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def euclidean(a, b):
    """metrics.rs:14 euclidean(): -squared-L2, AVX ordering iff len % 32 == 0."""
    a, b = _f32(a), _f32(b)
    return float(lib().orc_euclidean(_p(a, C.c_float), _p(b, C.c_float), a.size))


def sim_avx(a, b, portable=False):
    a, b = _f32(a), _f32(b)
    f = lib().orc_sim_avx_portable if portable else lib().orc_sim_avx
    return float(f(_p(a, C.c_float), _p(b, C.c_float), a.size))


def sim_scalar(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_sim_scalar(_p(a, C.c_float), _p(b, C.c_float), a.size))


def euclidean_batch(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty(a.shape[0], dtype=np.float32)
    lib().orc_euclidean_batch(_p(a, C.c_float), _p(b, C.c_float), a.shape[0], a.shape[1], _p(out, C.c_float))
    return out


def cut_ties(reset=False):
    """(select ties, eviction ties, order ties) since the last reset (process-wide): equal sims of different nodes on either
    side of a cut — the m-th / (m+1)-th candidate of select_neighbors, the two worst of w at an eviction — i.e. outcomes the
    reference leaves to BinaryHeap internals; and equal sims among the SELECTED candidates, where the set is pinned but the
    order of the tied pair in the adjacency list is not.  Parity fixtures are chosen so that the first is 0."""
    n = (int(lib().orc_cut_ties()), int(lib().orc_evict_ties()), int(lib().orc_order_ties()))
    if reset:
        lib().orc_cut_ties_reset()
    return n


def level_from_u(u, m):
    """core.rs:601-605 with the uniform draw injected."""
    return int(lib().orc_level_from_u(float(u), int(m)))


def draw_levels(n, m, seed):
    """Injected per-node levels: floor(-ln(u) * 1/ln(m)), u ~ U[0,1) f64 from a seeded generator."""
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    u = np.where(u == 0.0, np.nextafter(0.0, 1.0), u)
    lv = np.floor(-np.log(u) * (1.0 / math.log(m))).astype(np.int32)
    return lv


class HNSWError(Exception):
    pass


class Oracle:
    """Index<f32,f32> of the reference (core.rs:302-346) with u32 ids instead of names."""

    def __init__(self, dim, m=5, ef_construction=200):
        self.dim, self.m, self.ef_construction = int(dim), int(m), int(ef_construction)
        self._h = C.c_void_p(lib().orc_create(self.dim, self.m, self.ef_construction))

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- parameters / pub fields
    def params(self):
        out = np.zeros(8, dtype=np.int64)
        lib().orc_params(self._h, _p(out, C.c_int64))
        keys = ("data_dim", "m", "m_max", "m_max_0", "ef_construction", "node_count", "max_layer", "enterpoint")
        d = dict(zip(keys, (int(x) for x in out)))
        d["level_mult"] = float(lib().orc_level_mult(self._h))
        return d

    @property
    def node_count(self):
        return self.params()["node_count"]

    def n_ids(self):
        return int(lib().orc_n_ids(self._h))

    # -- mutation
    def add(self, vec, level, stats=False):
        v = _f32(vec)
        if v.size != self.dim:
            raise HNSWError("data dimension: %d does not match Index" % v.size)  # core.rs:390
        st = np.zeros(4, dtype=np.uint64)
        i = lib().orc_add(self._h, _p(v, C.c_float), v.size, int(level), _p(st, C.c_uint64) if stats else None)
        return (int(i), st) if stats else int(i)

    def add_batch(self, vecs, levels):
        v = _f32(vecs)
        lv = np.ascontiguousarray(levels, dtype=np.int32)
        assert v.ndim == 2 and v.shape[1] == self.dim and lv.size == v.shape[0]
        st = np.zeros(4, dtype=np.uint64)
        lib().orc_add_batch(self._h, v.shape[0], _p(v, C.c_float), _p(lv, C.c_int32), _p(st, C.c_uint64))
        return st

    def delete(self, node_id):
        if lib().orc_delete(self._h, int(node_id)) != 0:
            raise HNSWError("Node: %r does not exist" % (node_id,))  # core.rs:421

    def touched(self):
        n = int(lib().orc_touched(self._h, None, 0))
        out = np.empty(n, dtype=np.uint32)
        lib().orc_touched(self._h, _p(out, C.c_uint32), n)
        return out

    # -- search
    def search(self, q, k, ef=None, stats=False):
        q = _f32(q)
        if q.size != self.dim:
            raise HNSWError("data dimension: %d does not match Index" % q.size)  # core.rs:479
        ef = self.ef_construction if ef is None else int(ef)  # core.rs:485
        ids = np.empty(k, dtype=np.uint32)
        sims = np.empty(k, dtype=np.float32)
        st = np.zeros(4, dtype=np.uint64)
        n = lib().orc_search(self._h, _p(q, C.c_float), int(k), ef, _p(ids, C.c_uint32), _p(sims, C.c_float),
                             _p(st, C.c_uint64))
        return (ids[:n], sims[:n], st) if stats else (ids[:n], sims[:n])

    def search_batch(self, Q, k, ef=None, threads=1, stats=True):
        Q = _f32(Q)
        assert Q.ndim == 2 and Q.shape[1] == self.dim
        ef = self.ef_construction if ef is None else int(ef)
        nq = Q.shape[0]
        ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        sims = np.full((nq, k), -np.inf, dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        st = np.zeros((nq, 4), dtype=np.uint64) if stats else None
        secs = lib().orc_search_batch(self._h, nq, _p(Q, C.c_float), int(k), ef, _p(ids, C.c_uint32),
                                      _p(sims, C.c_float), _p(counts, C.c_uint32),
                                      _p(st, C.c_uint64) if stats else None, int(threads))
        return ids, sims, counts, st, float(secs)

    def search_level(self, q, ep, ef, level):
        q = _f32(q)
        ids = np.empty(ef, dtype=np.uint32)
        sims = np.empty(ef, dtype=np.float32)
        n = lib().orc_search_level(self._h, _p(q, C.c_float), int(ep), int(ef), int(level), _p(ids, C.c_uint32),
                                   _p(sims, C.c_float), int(ef))
        return ids[:n], sims[:n]

    # -- node getters
    def node_level(self, i):
        return int(lib().orc_node_level(self._h, int(i)))

    def node_n_levels(self, i):
        return int(lib().orc_node_n_levels(self._h, int(i)))

    def node_neighbors(self, i, level):
        n = int(lib().orc_node_neighbors(self._h, int(i), int(level), None, 0))
        out = np.empty(n, dtype=np.uint32)
        if n:
            lib().orc_node_neighbors(self._h, int(i), int(level), _p(out, C.c_uint32), n)
        return out

    def node_vector(self, i):
        out = np.empty(self.dim, dtype=np.float32)
        lib().orc_node_vector(self._h, int(i), _p(out, C.c_float))
        return out

    # -- flat graph exchange (see hnsw_oracle.cpp: rows are (node, level), level-major within a node)
    def export_graph(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().orc_graph_sizes(self._h, C.byref(a), C.byref(b), C.byref(c))
        n, rows, edges = a.value, b.value, c.value
        levels = np.empty(n, dtype=np.int32)
        offs = np.zeros(rows + 1, dtype=np.uint64)
        nbr = np.empty(max(edges, 1), dtype=np.uint32)
        entry, ml = C.c_int64(), C.c_int32()
        lib().orc_export(self._h, _p(levels, C.c_int32), _p(offs, C.c_uint64), _p(nbr, C.c_uint32), C.byref(entry),
                         C.byref(ml))
        return dict(n=n, levels=levels, row_offs=offs, nbrs=nbr[:edges], entry=int(entry.value), max_layer=int(ml.value))

    def import_graph(self, vecs, g):
        v = _f32(vecs)
        levels = np.ascontiguousarray(g["levels"], dtype=np.int32)
        offs = np.ascontiguousarray(g["row_offs"], dtype=np.uint64)
        nbr = np.ascontiguousarray(g["nbrs"], dtype=np.uint32)
        if nbr.size == 0:
            nbr = np.zeros(1, dtype=np.uint32)
        lib().orc_import(self._h, v.shape[0], _p(v, C.c_float), _p(levels, C.c_int32), _p(offs, C.c_uint64),
                         _p(nbr, C.c_uint32), int(g["entry"]), int(g["max_layer"]))
