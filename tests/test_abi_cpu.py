"""CPU-side checks of the drop-in boundary: the library loads, exports every declared symbol, and the host
layer behaves without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import redis_hnsw_b200 as r

    r.build()
    return r


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "hnsw_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hnsw_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = C.CDLL(built.SO_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libhnsw_b200.so does not export %s" % name
    from redis_hnsw_b200 import _lib

    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_no_cpu_fallback(built):
    """Without a CUDA device index creation must fail loudly (never silently compute on the CPU)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.HNSWError, match="no CUDA device"):
        built.DeviceIndex(128, 16, 200)
    with pytest.raises(built.HNSWError):
        built.l2_batch(np.zeros((1, 32), np.float32), np.zeros((1, 32), np.float32))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under redis_hnsw_b200/ may reference it."""
    pkg = os.path.join(ROOT, "redis_hnsw_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                assert "hnsw_oracle" not in src, f


def test_row_permutation_is_a_bijection():
    """The slab layout of distance.cuh (element i -> word ((c/V)*32 + t)*V + c%V) restated in numpy."""
    for dim in (32, 64, 96, 128, 256, 768, 1024):
        Cn = dim // 32
        V = 4 if Cn % 4 == 0 else (2 if Cn % 2 == 0 else 1)
        i = np.arange(dim)
        c, t = i >> 5, i & 31
        pos = ((c // V) * 32 + t) * V + (c % V)
        assert sorted(pos.tolist()) == list(range(dim))
        # lane t's V-wide vector g holds chunks g*V .. g*V+V-1 of residue t
        for g in range(Cn // V):
            for lane in (0, 7, 31):
                words = [(g * 32 + lane) * V + r for r in range(V)]
                elems = [int(i[pos == w][0]) for w in words]
                assert elems == [32 * (g * V + r) + lane for r in range(V)]


def test_level_draw_matches_reference_formula(oracle_mod):
    import math

    from redis_hnsw_b200 import data

    lv = data.draw_levels(10000, 16, seed=42)
    assert np.array_equal(lv, oracle_mod.draw_levels(10000, 16, 42))
    rng = np.random.default_rng(42)
    u = rng.random(10000)
    for k in (0, 1, 17, 9999):
        assert lv[k] == int(-math.log(u[k]) * (1.0 / math.log(16)))  # core.rs:601-605
        assert lv[k] == oracle_mod.level_from_u(u[k], 16)
    assert 0.05 < (lv >= 1).mean() < 0.075  # P(level >= 1) = 1/m


def test_datasets_are_seeded():
    from redis_hnsw_b200 import data

    a, qa = data.lowrank(1000, 128, seed=5, n_queries=10)
    b, qb = data.lowrank(1000, 128, seed=5, n_queries=10)
    assert np.array_equal(a, b) and np.array_equal(qa, qb) and a.dtype == np.float32
    gt = data.brute_force_topk(a, qa, 10)
    d = ((a[None, :, :] - qa[:, None, :]) ** 2).sum(-1)
    assert np.array_equal(np.sort(gt, 1), np.sort(np.argsort(d, 1)[:, :10], 1))
    assert data.recall_at_k(gt, gt) == 1.0


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/hnsw_b200.h must compile as C99 (no C++ or CUDA types in the signatures) and a C
    program must link against the library."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "hnsw_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { hnsw_index_t* h = 0; int rc = hnsw_index_create(4, 5, 16, -1, &h);\n'
                   '  printf("%d %s|%s\\n", rc, hnsw_version(), rc ? hnsw_last_error() : "");\n'
                   '  if (h) { hnsw_index_destroy(h); }\n  return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(root, "redis_hnsw_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lhnsw_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rc, rest = out.stdout.split(" ", 1)
    assert rc in ("0", "5")            # 5 = HNSW_ERR_CUDA on a machine without a device: loud failure, no fallback
    if rc == "5":
        assert "CUDA" in rest or "device" in rest


def test_bench_stdout_carries_only_the_json_line(tmp_path):
    """bench.py's contract is ONE JSON line on stdout; libraries that print banners on fd 1 (NCCL's version line at
    N > 1) must not end up next to it: after claim_stdout() fd 1 points at stderr and emit() writes to the real stdout."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "bench.claim_stdout()\n"
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')\n"      # what a C library does
            "print('a python print')\n"
            "bench.emit({'metric': 'x', 'value': 1})\n" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "x", "value": 1}
    assert "NCCL version" in out.stderr and "a python print" in out.stderr


def test_null_handle_and_bad_arguments_fail_cleanly(built):
    """"The library never aborts" (include/hnsw_b200.h): every entry point taking an index answers HNSW_ERR_INVALID for a
    NULL handle, create() rejects bad parameters before touching a device, and hnsw_last_error() carries the text."""
    from redis_hnsw_b200 import _lib

    L = _lib.lib()
    null = C.c_void_p(0)
    for name, (res, args) in _lib.SYMBOLS.items():
        if not args or args[0] is not _lib._vp or name == "hnsw_index_destroy":
            continue
        call_args = [null]
        for a in args[1:]:
            if a in (C.c_uint32, C.c_uint64, C.c_int, C.c_int32, C.c_int64):
                call_args.append(a(1))
            elif a is C.c_char_p:
                call_args.append(b"row_copy")
            else:
                call_args.append(None)      # NULL pointer
        assert getattr(L, name)(*call_args) == _lib.ERR_INVALID, name
        assert b"null index handle" in L.hnsw_last_error(), name
    L.hnsw_index_destroy(null)              # a no-op, like free(NULL)
    out = C.c_void_p(0)
    for dim, m, efc in ((0, 16, 200), (128, 0, 200), (128, 16, 0), (128, 16, 65537)):
        assert L.hnsw_index_create(dim, m, efc, -1, C.byref(out)) == _lib.ERR_INVALID
        assert not out.value
    assert L.hnsw_index_create(128, 16, 200, -1, None) == _lib.ERR_INVALID
    assert L.hnsw_l2_batch(None, None, 4, 32, None, -1) == _lib.ERR_INVALID
    assert L.hnsw_l2_batch(None, None, 0, 32, None, -1) == _lib.HNSW_OK       # nothing to do
    assert L.hnsw_version().startswith(b"hnsw_b200")
