"""The C++ host mirror of `VectorStore` (include/csgpu_store.hpp) over the C ABI: the reference's own test module
(src/vectordb/store.rs:826-1029) restated in C++ (tests/cpp/store_test.cpp), built by __graft_entry__.build()."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "store_test")


def _build():
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(
            os.path.getmtime(os.path.join(ROOT, "include", "csgpu_store.hpp")),
            os.path.getmtime(os.path.join(ROOT, "tests", "cpp", "store_test.cpp"))):
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(BIN)


def test_cpp_mirror_builds_and_fails_loudly_without_gpu():
    """CPU box: the mirror compiles and links against libcsgpu.so; opening a store must fail with the ABI's CUDA error
    (no CPU fallback). On a GPU box this test is meaningless and skipped."""
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([BIN, "--expect-no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "NO-GPU OK" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_tests():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr
    for name in ("test_vector_store_creation", "test_insert_and_search", "test_stats", "test_clear", "test_get_chunk",
                 "test_persistence", "test_guards", "test_additive_methods"):
        assert f"ok   {name}" in r.stdout
