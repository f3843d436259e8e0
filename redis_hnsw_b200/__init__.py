"""redis_hnsw_b200 — B200-native HNSW index + search engine behind the redis_hnsw operator surface.

The compute path is libhnsw_b200.so (hand-written sm_100a CUDA behind the C ABI of include/hnsw_b200.h);
this package only holds the host-side mirror of the reference's `Index` interface, the synthetic datasets and
the ctypes binding.  Importing never touches the GPU; creating an index does, and fails loudly without one.
"""
from . import data, sharding  # noqa: F401
from ._lib import BUILD_EXACT, BUILD_FAST, BUILD_SPEC, REDIS_MODULE_PATH, SO_PATH, build  # noqa: F401
from .index import DeviceIndex, HNSWError, Index, SearchResult, l2_batch, launch_count  # noqa: F401
