"""ctypes binding of libhnsw_b200.so (include/hnsw_b200.h).  There is no fallback: if the CUDA library is
missing or cannot be loaded, importing the compute API fails loudly."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libhnsw_b200.so")
REDIS_MODULE_PATH = os.path.join(_HERE, "libredis_hnsw_b200.so")  # redis-server --loadmodule <this>

HNSW_OK = 0
ERR_DIM_MISMATCH, ERR_EXISTS, ERR_NOT_FOUND, ERR_INVALID, ERR_CUDA, ERR_OOM = 1, 2, 3, 4, 5, 6
NO_NODE = 0xFFFFFFFF
BUILD_EXACT, BUILD_FAST, BUILD_SPEC = 0, 1, 2


class Params(C.Structure):
    _fields_ = [("data_dim", C.c_uint32), ("m", C.c_uint32), ("m_max", C.c_uint32), ("m_max_0", C.c_uint32),
                ("ef_construction", C.c_uint32), ("max_layer", C.c_int32), ("level_mult", C.c_double),
                ("node_count", C.c_uint64), ("n_ids", C.c_uint64), ("enterpoint", C.c_uint32), ("device", C.c_int32)]


class DeviceBuffer(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("bytes", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/hnsw_b200.h declares
_fp, _u32p, _u64p, _i32p, _i64p, _vp = (C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p)
SYMBOLS = {
    "hnsw_index_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(_vp)]),
    "hnsw_index_destroy": (None, [_vp]),
    "hnsw_index_reserve": (C.c_int, [_vp, C.c_uint64]),
    "hnsw_index_seed": (C.c_int, [_vp, C.c_uint64]),
    "hnsw_index_add": (C.c_int, [_vp, _fp, C.c_uint64, C.c_int32, _u32p]),
    "hnsw_index_add_batch": (C.c_int, [_vp, C.c_uint64, _fp, _i32p, C.c_int, _u32p]),
    "hnsw_index_touched": (C.c_int, [_vp, _u32p, C.c_uint64, _u64p]),
    "hnsw_index_delete": (C.c_int, [_vp, C.c_uint32]),
    "hnsw_index_search": (C.c_int, [_vp, _fp, C.c_uint64, C.c_uint32, C.c_uint32, _u32p, _fp, _u32p]),
    "hnsw_index_search_batch": (C.c_int, [_vp, C.c_uint64, _fp, C.c_uint32, C.c_uint32, _u32p, _fp, _u32p, _vp]),
    "hnsw_index_search_batch_device": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "hnsw_index_search_level": (C.c_int, [_vp, _fp, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, _fp, _u32p]),
    "hnsw_l2_batch": (C.c_int, [_fp, _fp, C.c_uint64, C.c_uint32, _fp, C.c_int]),
    "hnsw_index_params": (C.c_int, [_vp, C.POINTER(Params)]),
    "hnsw_index_node_level": (C.c_int, [_vp, C.c_uint32, _i32p]),
    "hnsw_index_node_neighbors": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _u32p, C.c_uint64, _u64p]),
    "hnsw_index_node_vector": (C.c_int, [_vp, C.c_uint32, _fp]),
    "hnsw_index_rows_batch": (C.c_int, [_vp, C.c_uint64, _u32p, _u32p, C.c_uint32, _u32p, _u32p]),
    "hnsw_index_graph_sizes": (C.c_int, [_vp, _u64p, _u64p, _u64p]),
    "hnsw_index_export_graph": (C.c_int, [_vp, _i32p, _u64p, _u32p, _i64p, _i32p]),
    "hnsw_index_export_vectors": (C.c_int, [_vp, _fp]),
    "hnsw_index_load_graph": (C.c_int, [_vp, C.c_uint64, _fp, _i32p, _u64p, _u32p, C.c_int64, C.c_int32]),
    "hnsw_index_device_buffers": (C.c_int, [_vp, C.POINTER(DeviceBuffer), C.c_uint32, _u32p]),
    "hnsw_index_replica_layout": (C.c_int, [_vp, _u64p]),
    "hnsw_index_prepare_replica": (C.c_int, [_vp, _u64p]),
    "hnsw_index_adopt_replica": (C.c_int, [_vp]),
    "hnsw_index_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "hnsw_launch_count": (C.c_uint64, []),
    "hnsw_index_build_stats": (C.c_int, [_vp, _u64p]),
    "hnsw_index_build_stats_ex": (C.c_int, [_vp, _u64p, C.c_uint32, _u32p]),
    "hnsw_last_error": (C.c_char_p, []),
    "hnsw_version": (C.c_char_p, []),
}

_lib = None


def _source_hash():
    """sha256 over everything the two libraries are compiled from (sources, headers, Makefiles)."""
    import hashlib

    h = hashlib.sha256()
    roots = [os.path.join(_HERE, "csrc"), os.path.join(os.path.dirname(_HERE), "include"),
             os.path.join(os.path.dirname(_HERE), "tests", "fake_redis")]
    for root in roots:
        for d, _, files in sorted(os.walk(root)):
            for f in sorted(files):
                if f.endswith((".cu", ".cuh", ".hpp", ".cpp", ".h", "Makefile")):
                    h.update(f.encode())
                    with open(os.path.join(d, f), "rb") as fh:
                        h.update(fh.read())
    return h.hexdigest()


def build(jobs=8, force=False):
    """Compile libhnsw_b200.so for sm_100a with the in-tree Makefile (nvcc cross-compiles without a GPU).

    The built libraries travel to the GPU box next to a stamp holding the hash of the sources they were made from; when
    the stamp matches, nothing is recompiled (file times do not survive every copy, and the object files under build/ do
    not travel)."""
    stamp = SO_PATH + ".stamp"
    want = _source_hash()
    host = os.path.join(os.path.dirname(_HERE), "tests", "fake_redis", "fake_redis_host")
    if not force and all(os.path.exists(p) for p in (SO_PATH, REDIS_MODULE_PATH, host, stamp)):
        with open(stamp) as fh:
            if fh.read().strip() == want:
                return SO_PATH
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j%d" % jobs], stdout=subprocess.DEVNULL)
    # the Redis module (host side, plain C++ over the C ABI) and the fake module host the tests drive it with
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc", "redis")], stdout=subprocess.DEVNULL)
    with open(stamp, "w") as fh:
        fh.write(want + "\n")
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("libhnsw_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C redis_hnsw_b200/csrc`); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib
