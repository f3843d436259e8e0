"""Host-side mirror of the reference's `hnsw::Index<f32,f32>` (src/hnsw/core.rs:302-346) over the C ABI.

Same operator names, argument meaning and error text as the reference: Index(name, dim, m, ef_construction)
<- Index::new (core.rs:322); add_node(name, data, update_fn) (core.rs:383); delete_node (core.rs:414);
search_knn(data, k) (core.rs:477).  Node names map to the dense device ids here, as the reference keeps them in
`nodes: HashMap<String, Node>` (core.rs:316).  All arithmetic runs in libhnsw_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib


class HNSWError(Exception):
    """HNSWError::{Str,String} of the reference (core.rs:24-46)."""

    def __init__(self, msg, code=_lib.ERR_INVALID):
        super().__init__(msg)
        self.code = code


def _check(rc):
    if rc != _lib.HNSW_OK:
        raise HNSWError(_lib.lib().hnsw_last_error().decode(), rc)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class SearchResult:
    """SearchResult {sim, name, data} (core.rs:48-62)."""
    __slots__ = ("sim", "name", "data")

    def __init__(self, sim, name, data):
        self.sim, self.name, self.data = sim, name, data

    def __repr__(self):
        return "sim: %r, name: %r" % (self.sim, self.name)


def l2_batch(a, b, device=-1):
    """euclidean() of metrics.rs:14 on row pairs, on the GPU, bit-identical to the reference's CPU paths."""
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape and a.ndim == 2
    out = np.empty(a.shape[0], np.float32)
    _check(_lib.lib().hnsw_l2_batch(_p(a, C.c_float), _p(b, C.c_float), a.shape[0], a.shape[1], _p(out, C.c_float), device))
    return out


def launch_count():
    return int(_lib.lib().hnsw_launch_count())


class DeviceIndex:
    """Thin id-based wrapper of the C ABI (what the reference's Rust host would bind)."""

    def __init__(self, dim, m=5, ef_construction=200, device=-1):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        _check(self._L.hnsw_index_create(int(dim), int(m), int(ef_construction), int(device), C.byref(self._h)))
        self.dim = int(dim)

    def close(self):
        if getattr(self, "_h", None):
            self._L.hnsw_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters
    def params(self):
        p = _lib.Params()
        _check(self._L.hnsw_index_params(self._h, C.byref(p)))
        d = {k: getattr(p, k) for k, _ in _lib.Params._fields_}
        d["enterpoint"] = -1 if d["enterpoint"] == _lib.NO_NODE else d["enterpoint"]
        return d

    def reserve(self, n):
        _check(self._L.hnsw_index_reserve(self._h, int(n)))

    def seed(self, s):
        _check(self._L.hnsw_index_seed(self._h, int(s)))

    def set_option(self, name, value):
        _check(self._L.hnsw_index_set_option(self._h, name.encode(), int(value)))

    # -- mutation
    def add(self, vec, level=-1):
        v = _f32(vec).ravel()
        out = C.c_uint32()
        _check(self._L.hnsw_index_add(self._h, _p(v, C.c_float), v.size, int(level), C.byref(out)))
        return int(out.value)

    def add_batch(self, vecs, levels=None, mode=_lib.BUILD_EXACT):
        v = _f32(vecs)
        assert v.ndim == 2
        if v.shape[1] != self.dim:
            raise HNSWError("data dimension: %d does not match Index" % v.shape[1], _lib.ERR_DIM_MISMATCH)
        lv = None if levels is None else np.ascontiguousarray(levels, dtype=np.int32)
        first = C.c_uint32()
        _check(self._L.hnsw_index_add_batch(self._h, v.shape[0], _p(v, C.c_float),
                                            None if lv is None else _p(lv, C.c_int32), int(mode), C.byref(first)))
        return int(first.value)

    def delete(self, node_id):
        _check(self._L.hnsw_index_delete(self._h, int(node_id)))

    def touched(self):
        n = C.c_uint64()
        _check(self._L.hnsw_index_touched(self._h, None, 0, C.byref(n)))
        out = np.empty(n.value, np.uint32)
        if n.value:
            _check(self._L.hnsw_index_touched(self._h, _p(out, C.c_uint32), n.value, C.byref(n)))
        return out

    def build_stats(self):
        out = np.zeros(4, np.uint64)
        _check(self._L.hnsw_index_build_stats(self._h, _p(out, C.c_uint64)))
        d = dict(inserts=int(out[0]), conflicts=int(out[1]), reprunes=int(out[2]), dist_evals=int(out[3]))
        ex = np.zeros(16, np.uint64)
        n = C.c_uint32()
        _check(self._L.hnsw_index_build_stats_ex(self._h, _p(ex, C.c_uint64), ex.size, C.byref(n)))
        names = ("fast_worklist_dropped", "fast_reprunes_skipped", "fast_edges_refused", "spec_rounds", "spec_executions",
                 "spec_dist_evals_wasted", "spec_exact_fallbacks", "spec_max_window", "spec_rows_as_operations")
        d.update({k: int(ex[4 + i]) for i, k in enumerate(names)})
        return d

    # -- search
    def search(self, q, k, ef=0):
        q = _f32(q).ravel()
        ids = np.empty(k, np.uint32)
        sims = np.empty(k, np.float32)
        n = C.c_uint32()
        _check(self._L.hnsw_index_search(self._h, _p(q, C.c_float), q.size, int(k), int(ef), _p(ids, C.c_uint32),
                                         _p(sims, C.c_float), C.byref(n)))
        return ids[:n.value], sims[:n.value]

    def search_batch(self, Q, k, ef=0, stats=False, out=None):
        """Q: [nq, dim] host array (pinned or pageable).  Returns ids [nq,k], sims [nq,k], counts [nq] (+ stats [nq,4])."""
        Q = _f32(Q)
        assert Q.ndim == 2
        if Q.shape[1] != self.dim:
            raise HNSWError("data dimension: %d does not match Index" % Q.shape[1], _lib.ERR_DIM_MISMATCH)
        nq = Q.shape[0]
        if out is None:
            ids = np.empty((nq, k), np.uint32)
            sims = np.empty((nq, k), np.float32)
            counts = np.empty(nq, np.uint32)
        else:
            ids, sims, counts = out
        st = np.zeros((nq, 4), np.uint32) if stats else None
        _check(self._L.hnsw_index_search_batch(self._h, nq, _p(Q, C.c_float), int(k), int(ef), _p(ids, C.c_uint32),
                                               _p(sims, C.c_float), _p(counts, C.c_uint32),
                                               None if st is None else st.ctypes.data_as(C.c_void_p)))
        return (ids, sims, counts, st) if stats else (ids, sims, counts)

    def search_batch_device(self, nq, d_queries, k, ef, d_ids, d_sims, d_counts, d_stats=0, stream=0):
        """All arguments are raw device pointers (ints); enqueues on `stream` without synchronising."""
        _check(self._L.hnsw_index_search_batch_device(self._h, int(nq), d_queries, int(k), int(ef), d_ids, d_sims,
                                                      d_counts, d_stats or None, stream or None))

    def search_level(self, q, ep, ef, level):
        q = _f32(q).ravel()
        ids = np.empty(ef, np.uint32)
        sims = np.empty(ef, np.float32)
        n = C.c_uint32()
        _check(self._L.hnsw_index_search_level(self._h, _p(q, C.c_float), int(ep), int(ef), int(level),
                                               _p(ids, C.c_uint32), _p(sims, C.c_float), C.byref(n)))
        return ids[:n.value], sims[:n.value]

    # -- node getters
    def node_level(self, i):
        lv = C.c_int32()
        _check(self._L.hnsw_index_node_level(self._h, int(i), C.byref(lv)))
        return int(lv.value)

    def node_neighbors(self, i, level):
        n = C.c_uint64()
        buf = np.empty(512, np.uint32)
        _check(self._L.hnsw_index_node_neighbors(self._h, int(i), int(level), _p(buf, C.c_uint32), buf.size, C.byref(n)))
        if n.value > buf.size:
            buf = np.empty(n.value, np.uint32)
            _check(self._L.hnsw_index_node_neighbors(self._h, int(i), int(level), _p(buf, C.c_uint32), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def node_vector(self, i):
        out = np.empty(self.dim, np.float32)
        _check(self._L.hnsw_index_node_vector(self._h, int(i), _p(out, C.c_float)))
        return out

    # -- whole graph
    def export_graph(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._L.hnsw_index_graph_sizes(self._h, C.byref(a), C.byref(b), C.byref(c)))
        n, rows, edges = a.value, b.value, c.value
        levels = np.empty(max(n, 1), np.int32)
        offs = np.zeros(rows + 1, np.uint64)
        nbrs = np.empty(max(edges, 1), np.uint32)
        entry, ml = C.c_int64(), C.c_int32()
        _check(self._L.hnsw_index_export_graph(self._h, _p(levels, C.c_int32), _p(offs, C.c_uint64), _p(nbrs, C.c_uint32),
                                               C.byref(entry), C.byref(ml)))
        return dict(n=n, levels=levels[:n], row_offs=offs, nbrs=nbrs[:edges], entry=int(entry.value),
                    max_layer=int(ml.value))

    def export_vectors(self):
        n = self.params()["n_ids"]
        out = np.empty((n, self.dim), np.float32)
        if n:
            _check(self._L.hnsw_index_export_vectors(self._h, _p(out, C.c_float)))
        return out

    def load_graph(self, vecs, g):
        v = _f32(vecs)
        levels = np.ascontiguousarray(g["levels"], np.int32)
        offs = np.ascontiguousarray(g["row_offs"], np.uint64)
        nbrs = np.ascontiguousarray(g["nbrs"], np.uint32)
        if nbrs.size == 0:
            nbrs = np.zeros(1, np.uint32)
        _check(self._L.hnsw_index_load_graph(self._h, v.shape[0], _p(v, C.c_float), _p(levels, C.c_int32),
                                             _p(offs, C.c_uint64), _p(nbrs, C.c_uint32), int(g["entry"]),
                                             int(g["max_layer"])))

    # -- replication
    def device_buffers(self):
        n = C.c_uint32()
        arr = (_lib.DeviceBuffer * 16)()
        _check(self._L.hnsw_index_device_buffers(self._h, arr, 16, C.byref(n)))
        return [(int(arr[i].ptr or 0), int(arr[i].bytes)) for i in range(n.value)]

    def replica_layout(self):
        out = np.zeros(8, np.uint64)
        _check(self._L.hnsw_index_replica_layout(self._h, _p(out, C.c_uint64)))
        return out

    def prepare_replica(self, layout):
        lay = np.ascontiguousarray(layout, np.uint64)
        _check(self._L.hnsw_index_prepare_replica(self._h, _p(lay, C.c_uint64)))

    def adopt_replica(self):
        _check(self._L.hnsw_index_adopt_replica(self._h))


class Index:
    """`Index<f32,f32>` as the reference's command handlers use it (names in, names out)."""

    def __init__(self, name, data_dim, m=5, ef_construction=200, device=-1):
        self.name = name                      # core.rs:304
        self.mfunc_kind = "Euclidean"         # core.rs:306,332
        self._dev = DeviceIndex(data_dim, m, ef_construction, device)
        self._ids = {}                        # name -> id   (core.rs:316 `nodes`)
        self._names = []                      # id -> name (None once deleted)

    # pub fields (core.rs:303-319)
    def __getattr__(self, key):
        if key in ("data_dim", "m", "m_max", "m_max_0", "ef_construction", "level_mult", "node_count", "max_layer"):
            return self._dev.params()[key]
        raise AttributeError(key)

    @property
    def enterpoint(self):
        e = self._dev.params()["enterpoint"]
        return None if e < 0 else self._names[e]

    @property
    def nodes(self):
        return self._ids

    @property
    def device_index(self):
        return self._dev

    def add_node(self, name, data, update_fn=None, level=-1):
        """core.rs:383-412.  `update_fn(name, node_id)` is called for every node whose adjacency changed
        (core.rs:580-584).  `level` injects the level draw (tests); -1 = draw."""
        data = _f32(data).ravel()
        if data.size != self._dev.dim:
            raise HNSWError("data dimension: %d does not match Index" % data.size, _lib.ERR_DIM_MISMATCH)  # :390
        if self._dev.params()["node_count"] > 0 and name in self._ids:
            raise HNSWError("Node: %r already exists" % name, _lib.ERR_EXISTS)  # :407-409
        try:
            nid = self._dev.add(data, level)
        except HNSWError:
            # a failed add still consumed its id (a tombstone on the device): keep the name table aligned
            while len(self._names) < self._dev.params()["n_ids"]:
                self._names.append(None)
            raise
        assert nid == len(self._names)
        self._ids[name] = nid
        self._names.append(name)
        if update_fn is not None:
            for t in self._dev.touched():
                update_fn(self._names[int(t)], int(t))

    def delete_node(self, name, update_fn=None):
        """core.rs:414-475."""
        if name not in self._ids:
            raise HNSWError("Node: %r does not exist" % name, _lib.ERR_NOT_FOUND)  # :421
        nid = self._ids.pop(name)
        self._dev.delete(nid)
        self._names[nid] = None
        if update_fn is not None:
            for t in self._dev.touched():
                update_fn(self._names[int(t)], int(t))

    def search_knn(self, data, k, ef=0):
        """core.rs:477-486: ef = ef_construction unless the `ef` extension is given.  Result names are the last
        '.'-segment of the node name (core.rs:885-887); data is a copy of the stored vector (core.rs:888)."""
        data = _f32(data).ravel()
        if data.size != self._dev.dim:
            raise HNSWError("data dimension: %d does not match Index" % data.size, _lib.ERR_DIM_MISMATCH)  # :479
        ids, sims = self._dev.search(data, k, ef)
        return [SearchResult(float(s), self._names[int(i)].split(".")[-1], self._dev.node_vector(int(i)))
                for i, s in zip(ids, sims) if int(i) < len(self._names) and self._names[int(i)] is not None]
