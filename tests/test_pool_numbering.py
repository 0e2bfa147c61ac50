"""Path numbering of the slot-stable pool (csrc/rtb_device.cuh: chunk_path), checked on the host build of the device
header: every camera path number below `total` is started by exactly one chunk, each chunk's sequence is strictly
increasing (so "my next number is >= total" means the chunk is finished for good), and consecutive numbers of a chunk
stay inside one 32-path block (= one 8x4 pixel tile, main.rs:751-754 sample loop order is irrelevant to the image)."""
import ctypes as C

import numpy as np
import pytest

import helpers as H


@pytest.fixture(scope="module")
def lib():
    lib = H.build_emul()
    lib.emul_chunk_path.restype = C.c_uint64
    lib.emul_chunk_path.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    return lib


@pytest.mark.parametrize("n_chunks,total", [(1, 1), (1, 1000), (4, 100), (7, 32 * 7 * 5), (20, 5000), (20, 4999), (33, 70000)])
def test_every_path_number_is_started_exactly_once(lib, n_chunks, total):
    seen = np.zeros(total, dtype=np.int32)
    for c in range(n_chunks):
        prev = -1
        m = 0
        while True:
            p = lib.emul_chunk_path(m, c, n_chunks)
            assert p > prev  # strictly increasing: once >= total, always >= total
            prev = p
            if p >= total:
                break
            seen[p] += 1
            m += 1
        # the numbers after the first one past the end are past the end as well
        assert lib.emul_chunk_path(m + 1, c, n_chunks) >= total and lib.emul_chunk_path(m + 40, c, n_chunks) >= total
    assert (seen == 1).all()


def test_blocks_of_32_consecutive_numbers(lib):
    n_chunks = 11
    for c in (0, 5, 10):
        for k in range(4):
            block = [lib.emul_chunk_path(32 * k + j, c, n_chunks) for j in range(32)]
            assert block == list(range(block[0], block[0] + 32)) and block[0] % 32 == 0
            assert block[0] // 32 == k * n_chunks + c
