"""The C-ABI library loads on a CPU-only box and exports every symbol include/csgpu.h declares.
No compute calls here (there is no GPU); csgpu_create must fail loudly, not fall back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "csgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csgpu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from codesearch_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"libcsgpu.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} is declared in csgpu.h but not bound in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.csgpu_abi_version() == 8


def test_stats_struct_layout_matches_header():
    from codesearch_b200 import _lib
    # 6 u64 + 4 u32 + f32 + u32 + 8 u64
    assert ctypes.sizeof(_lib.Stats) == 6 * 8 + 4 * 4 + 4 + 4 + 8 * 8 + 4 * 8 + 5 * 8 + 4 + 4   # ... + coalesced_passes, coalesced_queries, prefilter_rescored, shadow_bytes + 5 byte-prefilter counters + batch_route, filter_max_err


def test_decode_keys_is_pure_host():
    import numpy as np
    from codesearch_b200 import _lib
    lib = _lib.load()

    def okey(f):
        u = int(np.float32(f).view(np.uint32))
        return u ^ (0xFFFFFFFF if u >> 31 else 0x80000000)
    keys = np.array([(okey(-1e-8) << 32) | 5, (okey(0.0) << 32) | 2, (okey(0.25) << 32) | 0xFFFFFFFE,
                     _lib.KEY_EMPTY], dtype=np.uint64)
    assert list(keys) == sorted(keys)   # order-preserving encoding
    ids = np.zeros(4, np.uint32); dist = np.zeros(4, np.float32); n = ctypes.c_uint32()
    lib.csgpu_decode_keys(keys.ctypes.data_as(_lib._u64p), 4, ids.ctypes.data_as(_lib._u32p),
                          dist.ctypes.data_as(_lib._f32p), ctypes.byref(n))
    assert n.value == 3
    assert ids[:3].tolist() == [5, 2, 0xFFFFFFFE]
    assert dist[:3].tolist() == [np.float32(-1e-8), 0.0, 0.25]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from codesearch_b200 import _lib, VectorStore, CsgpuError
    with pytest.raises(CsgpuError) as e:
        VectorStore.new(None, 384)
    assert e.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_argument_errors_need_no_device():
    from codesearch_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.csgpu_create(ctypes.byref(h), 0, 0, None, 1) == _lib.ERR_ARG
    assert lib.csgpu_create(ctypes.byref(h), 384, 7, None, 1) == _lib.ERR_ARG
    assert lib.csgpu_create(ctypes.byref(h), 384, 0, None, 9) == _lib.ERR_ARG
    assert lib.csgpu_search(None, None, 0, 0, None, None, None) == _lib.ERR_ARG
    assert b"null index" in lib.csgpu_last_error()


def test_product_never_imports_oracle():
    """The product package must not route through oracle/ (that would void parity)."""
    pkg = os.path.join(ROOT, "codesearch_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_encode_decode_keys_round_trip():
    import numpy as np
    from codesearch_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 2**32, size=37, dtype=np.uint64).astype(np.uint32)
    dist = np.sort(rng.random(37).astype(np.float32))
    keys = np.zeros(50, np.uint64)
    lib.csgpu_encode_keys(ids.ctypes.data_as(_lib._u32p), dist.ctypes.data_as(_lib._f32p), 37, 50, keys.ctypes.data_as(_lib._u64p))
    assert (keys[37:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    assert (np.diff(keys[:37].astype(np.float64)) >= 0).all()          # ascending distance <=> ascending key
    oi = np.zeros(50, np.uint32); od = np.zeros(50, np.float32); n = ctypes.c_uint32()
    lib.csgpu_decode_keys(keys.ctypes.data_as(_lib._u64p), 50, oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p), ctypes.byref(n))
    assert n.value == 37 and np.array_equal(oi[:37], ids) and np.array_equal(od[:37], dist)


def test_header_is_valid_c_and_the_c_host_compiles():
    """include/csgpu.h is a C header (the reference's host is Rust over a C ABI): the plain-C latency bench that uses most of
    it must compile as C11 with warnings as errors, and the header alone as C99 (no C++-isms, no torch types)."""
    import os
    import shutil
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    inc = os.path.join(root, "include")
    subprocess.check_call([gcc, "-std=c11", "-Wall", "-Werror", "-D_POSIX_C_SOURCE=200809L", "-fsyntax-only", "-I", inc,
                           os.path.join(root, "tools", "bench_c_abi.c")])
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "only_header.c")
        with open(src, "w") as f:
            f.write('#include "csgpu.h"\nint main(void) { csgpu_stats_t s; (void)s; return (int)(CSGPU_ABI_VERSION == 0); }\n')
        subprocess.check_call([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, src])
