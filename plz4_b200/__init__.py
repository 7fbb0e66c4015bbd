"""plz4_b200 — B200-native LZ4 block engine behind plz4's independent-block path.

The product is plz4_b200/libplz4cu.so (CUDA sm_100a kernels + the C ABI of include/plz4cu.h);
this package is the thin host-side mirror of the reference interface used by tests and bench.
"""
from . import _lib  # noqa: F401
from .api import (  # noqa: F401
    Dict, Lz4Error, Plz4cuError, compress_batch, compress_block, compress_block_bound, compress_frame_device,
    decompress_batch, decompress_block, decompress_frame_device, device_count, init, init_devices, lz4_corrupted,
)
from .stream import NewReader, NewWriter, Reader, StreamError, Writer, write_skip_frame_header  # noqa: F401,E402
