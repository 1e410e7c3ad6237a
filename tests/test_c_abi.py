"""The C-ABI library loads and exports every symbol include/sangnom_cuda.h declares; struct layouts in the
ctypes binding match the header; without a GPU every compute entry fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from pysangnom import cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "sangnom_cuda.h")).read()


def declared_functions():
    return sorted(set(re.findall(r"SN_API\s+[\w\s\*]+?\b(sangnom_cuda_\w+)\s*\(", HEADER)))


def test_every_declared_symbol_is_exported():
    lib = cuda.load()
    names = declared_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sangnom_cuda.h but not exported"
    assert sorted(cuda.EXPORTS) == names


def test_abi_version_and_struct_layout():
    lib = cuda.load()
    assert lib.sangnom_cuda_abi_version() == int(re.search(r"#define SANGNOM_CUDA_ABI_VERSION (\d+)", HEADER).group(1))
    # sn_plane_job: 2 x (ptr, ptrdiff) + 4 ints + float + 2 ints, 8-byte aligned
    assert C.sizeof(cuda.SnPlaneJob) == 64
    # sn_config: 7 ints, pad, u64 device_mask, int copy_threads, pad
    assert C.sizeof(cuda.SnConfig) == 48 and cuda.SnConfig.device_mask.offset == 32 and cuda.SnConfig.copy_threads.offset == 40
    assert C.sizeof(cuda.SnStats) == 48
    assert cuda.SnPlaneJob.threshold.offset == 48 and cuda.SnPlaneJob.frame.offset == 56


def test_threshold_scaling_matches_reference_formula():
    assert int(cuda.threshold(48, 8, 1)) == 63 and int(cuda.threshold(48, 10, 2)) == 252
    assert int(cuda.threshold(47, 16, 2)) == 15792 and cuda.threshold(24, 32, 4) == 0.123046875


def test_offset_resolution():
    assert cuda.resolve_offset(1, False) == 0 and cuda.resolve_offset(2, True) == 1
    assert cuda.resolve_offset(0, True) == 0 and cuda.resolve_offset(0, False) == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cuda.SangNomCudaError) as e:
        cuda.Context(1, 1920, 1080)
    assert e.value.code == cuda.SN_ERR_CUDA and "no CPU path" in str(e.value)


def test_create_rejects_bad_config():
    lib = cuda.load()
    h = C.c_void_p()
    bad = cuda.SnConfig(cuda.ABI_VERSION + 1, 0, 1, 64, 64, 0, 0)
    assert lib.sangnom_cuda_create(C.byref(bad), C.byref(h)) == cuda.SN_ERR_INVALID
    bad = cuda.SnConfig(cuda.ABI_VERSION, 0, 3, 64, 64, 0, 0)
    assert lib.sangnom_cuda_create(C.byref(bad), C.byref(h)) == cuda.SN_ERR_INVALID
    assert b"sample_type" in lib.sangnom_cuda_last_error(None)
