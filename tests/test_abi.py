"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mvslam_b200.h declares, struct layouts agree, host-only entry points work, and the product
refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import mvslam_b200 as mvs
from oracle import cbind as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mvslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = mvs.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mvslam_b200.h but not exported"
    assert L.mvs_abi_version() == 3


def test_struct_layouts():
    assert ctypes.sizeof(mvs.PairResult) == mvs.RESULT_DTYPE.itemsize == 376
    assert mvs.MATCH_DTYPE.itemsize == 12
    assert ctypes.sizeof(mvs.RansacParams) == 40 and ctypes.sizeof(mvs.MatchParams) == 24


def test_host_sample_table_matches_oracle():
    for seed, pid, n, H in [(0, 0, 8, 16), (1, 5, 91, 1024), (2**63, 2**40, 8192, 300)]:
        assert np.array_equal(mvs.sample_table(seed, pid, n, H), orc.sample_table(seed, pid, n, H))


def test_host_pnp_sample_table_matches_oracle():
    for seed, pid, n, H in [(0, 0, 4, 8), (1, 5, 91, 100), (2**63, 2**40, 5000, 300)]:
        t = mvs.pnp_sample_table(seed, pid, n, H)
        assert np.array_equal(t, orc.pnp_sample_table(seed, pid, n, H))
        assert t.max() < n and all(len(set(r)) == 4 for r in t.tolist())
    assert mvs.pnp_sample_table(3, 1, 50, 4)[0].tolist() == [0, 1, 2, 3]


def test_struct_layouts_of_the_added_stages():
    assert mvs.KEYPOINT_DTYPE.itemsize == 24 and ctypes.sizeof(mvs.OrbParams) == 16
    assert mvs.PNP_RESULT_DTYPE.itemsize == 208 and ctypes.sizeof(mvs.PnpParams) == 40
    assert mvs.BA_OBS_DTYPE.itemsize == 48 and mvs.BA_RESULT_DTYPE.itemsize == 24 and ctypes.sizeof(mvs.BaParams) == 32
    assert len(mvs.STAGES) == 16


def test_status_strings():
    L = mvs.load_library()
    assert L.mvs_status_string(0) == b"ok"
    assert b"CUDA" in L.mvs_status_string(mvs.E_CUDA)


def test_host_alloc_needs_a_device_and_never_aborts():
    """mvs_host_alloc hands out page-locked memory through the CUDA runtime: NULL (not a crash) without a device,
    a usable buffer with one; mvs_host_free(NULL) is a no-op."""
    import torch
    L = mvs.load_library()
    L.mvs_host_alloc.restype = ctypes.c_void_p
    L.mvs_host_alloc.argtypes = [ctypes.c_size_t]
    L.mvs_host_free.restype = None
    L.mvs_host_free.argtypes = [ctypes.c_void_p]
    L.mvs_host_free(None)
    p = L.mvs_host_alloc(4096)
    if torch.cuda.is_available():
        assert p
        ctypes.memset(p, 0x5A, 4096)
        assert ctypes.string_at(p, 4) == b"ZZZZ"
        L.mvs_host_free(p)
    else:
        assert not p


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(mvs.MvsError) as e:
        mvs.Context(0)
    assert e.value.status == mvs.E_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mvslam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "").replace("oracle's", ""), f


def test_cpp_adapters_compile_and_refuse_to_run_without_gpu():
    import subprocess
    import torch
    exe = os.path.join(ROOT, "tests", "cpp", "test_adapters")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_adapters.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "mvslam_b200"), "-lmvslam_b200", "-Wl,-rpath,$ORIGIN/../../mvslam_b200"],
                   check=True)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stdout


def test_tensor_matcher_sass_uses_tcgen05_tma_and_tmem():
    """Static evidence (no GPU): the shipped library's knn2_hamming_tc_kernel issues integer tcgen05 MMAs (UTCIMMA),
    TMA tensor loads (UTMALDG), packed TMEM loads (LDTM ... PACK16BIT) and packed 16-bit min/max (VIMNMX.S16x2)."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "mvslam_b200", "libmvslam_b200.so")
    out = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    start = out.find("knn2_hamming_tc_kernel")
    assert start >= 0, "kernel not found in the library"
    end = out.find("Function :", start)
    body = out[start:end if end > 0 else None]
    for mnemonic in ("UTCIMMA", "UTMALDG", "LDTM", "PACK16BIT", "VIMNMX.S16x2", "UTCBAR"):
        assert mnemonic in body, f"{mnemonic} missing from knn2_hamming_tc_kernel"
    assert "sm_100a" in out or "sm_100" in out


def test_shard_bounds_host_function_matches_python():
    """mvs_shard_bounds (csrc/sharded.cu) is pure host code: same contiguous slices as mvslam_b200.shard.shard_bounds"""
    import ctypes
    from mvslam_b200 import shard
    L = mvs.load_library()
    for n in (0, 1, 7, 8, 130816):
        for world in (1, 2, 3, 8):
            covered = 0
            for r in range(world):
                lo, hi = ctypes.c_int64(), ctypes.c_int64()
                L.mvs_shard_bounds(ctypes.c_int64(n), world, r, ctypes.byref(lo), ctypes.byref(hi))
                assert (lo.value, hi.value) == shard.shard_bounds(n, world, r)
                assert lo.value == covered
                covered = hi.value
            assert covered == n
