"""The host copy pool of libsangnom_cuda (csrc/host_copy_pool.h: worker threads that pack / unpack pageable frames) on
its own, without a GPU: random batches of strided row copies against memcpy, one to sixteen threads, small batches
(single-threaded short cut) and multi-megabyte ones (worker path), many rounds through one pool object."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-f", os.path.join(HERE, "emul", "Makefile")], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "emul", "libkernel_emul.so"))
    L.emul_copy_pool_selftest.restype = C.c_int
    L.emul_copy_pool_selftest.argtypes = [C.c_int, C.c_int, C.c_uint]
    return L


@pytest.mark.parametrize("threads", [1, 2, 3, 8, 16])
def test_copy_pool_matches_memcpy(emul, threads):
    assert emul.emul_copy_pool_selftest(threads, 24, 1234 + threads) == 0
