"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/panman_b200.h declares; without a device it refuses to compute (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import panman_b200 as pb

    pb.build_library()
    return pb.load_library()


def test_exports_match_header(lib):
    header = open(os.path.join(ROOT, "include", "panman_b200.h")).read()
    declared = set(re.findall(r"\b(pmb_[a-z_]+)\s*\(", header))
    assert {"pmb_create", "pmb_set_tree", "pmb_run_nuc", "pmb_upload_nuc", "pmb_run_resident", "pmb_download",
            "pmb_destroy", "pmb_last_error"} <= declared
    raw = C.CDLL(os.path.join(ROOT, "panman_b200", "libpanman_b200.so"))
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/panman_b200.h but not exported"
    from panman_b200.lib import EXPORTS

    assert set(EXPORTS) == declared


def test_host_library_exports_match_header(lib):
    header = open(os.path.join(ROOT, "include", "panman_b200_host.h")).read()
    declared = set(re.findall(r"\b(pmh_[a-z_0-9]+)\s*\(", header))
    assert {"pmh_tree_from_newick", "pmh_build_from_msa", "pmh_pangraph_load", "pmh_pangraph_run"} <= declared
    raw = C.CDLL(os.path.join(ROOT, "panman_b200", "libpanman_b200_host.so"))
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/panman_b200_host.h but not exported"


def test_version_and_null_safety(lib):
    assert b"sm_100a" in lib.pmb_version()
    assert lib.pmb_last_error(None) == b"null context"
    assert lib.pmb_set_option(None, b"x", 1) == -1


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for GPU-less hosts")
    import panman_b200 as pb

    with pytest.raises(pb.PanmanError) as e:
        pb.Context(0)
    assert e.value.code == -2  # PMB_ERR_CUDA
    # the raw ABI keeps the context alive for the message, and every compute entry refuses
    h = C.c_void_p()
    assert lib.pmb_create(C.byref(h), 0) == -2
    assert b"CUDA" in lib.pmb_last_error(h)
    import numpy as np

    off = np.asarray([0, 2, 2, 2], np.int32)
    idx = np.asarray([1, 2], np.int32)
    row = np.asarray([-1, 0, 1], np.int32)
    assert lib.pmb_set_tree(h, 3, 0, off.ctypes.data, idx.ctypes.data, row.ctypes.data) == -2
    assert lib.pmb_run_resident(h, 0, 0) == -2
    lib.pmb_destroy(h)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under panman_b200/ or include/ may reference it."""
    bad = []
    for base in ("panman_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|fs_oracle|libpanman_ref|emul_kernels|libemul", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_group_entries_without_a_device(lib):
    """The multi-GPU entries behave like the single-context ones on a GPU-less host: creation reports PMB_ERR_CUDA with a
    message, nothing computes, nothing crashes; the column ranges need no device at all."""
    import torch

    a, b = C.c_int64(), C.c_int64()
    assert lib.pmb_group_column_range(8, 5_000_000, 7, C.byref(a), C.byref(b)) == 0 and (a.value, b.value) == (4375552, 5_000_000)
    assert lib.pmb_group_column_range(0, 10, 0, C.byref(a), C.byref(b)) == -1
    assert lib.pmb_group_column_range(4, 10, 4, C.byref(a), C.byref(b)) == -1
    assert lib.pmb_group_last_error(None) == b"null group"
    assert lib.pmb_group_world(None) == 0
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for GPU-less hosts")
    import panman_b200 as pb

    with pytest.raises(pb.PanmanError) as e:
        pb.Group([0, 1])
    assert e.value.code == -2 and "CUDA" in str(e.value)
    h = C.c_void_p()
    dev = (C.c_int * 2)(0, 1)
    assert lib.pmb_group_create(C.byref(h), dev, 2, 0, 1) == -1  # more local ranks than the world holds
    assert lib.pmb_group_create(C.byref(h), dev, 2, 0, 2) == -2
    assert b"rank 0" in lib.pmb_group_last_error(h)
    assert lib.pmb_group_run_async(h, 0, 0) == -5   # PMB_ERR_NO_INPUT: nothing was uploaded
    assert lib.pmb_group_wait(h) == -2
    lib.pmb_group_destroy(h)
