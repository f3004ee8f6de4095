#!/usr/bin/env python
"""Generates the `extern "C"` block of integration/rust/gpu.rs from include/csgpu.h, so the Rust binding a maintainer
pastes into flupkede/codesearch (src/vectordb/gpu.rs) can never drift from the header: one declaration per exported symbol,
same order, same arity, same integer widths.

  python tools/gen_rust_ffi.py            # rewrites the block between the GENERATED markers in integration/rust/gpu.rs
  python tools/gen_rust_ffi.py --check    # exit 1 if the file is out of date (tests/test_rust_binding.py runs this logic)

The parser is shared with the test: parse_header() -> [(name, ret, [(ctype, argname), ...])].
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "csgpu.h")
GPU_RS = os.path.join(ROOT, "integration", "rust", "gpu.rs")
BEGIN = "    // ---- GENERATED from include/csgpu.h by tools/gen_rust_ffi.py: do not edit by hand ----"
END = "    // ---- END GENERATED ----"

# C type (whitespace-normalised, pointer stars attached) -> Rust type
CTYPE_TO_RUST = {
    "int": "c_int", "void": "()", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "float": "f32",
    "const char*": "*const c_char", "void*": "*mut c_void", "const void*": "*const c_void",
    "csgpu_index**": "*mut *mut CsgpuIndex", "csgpu_index*": "*mut CsgpuIndex", "const csgpu_index*": "*const CsgpuIndex",
    "csgpu_index*const*": "*const *mut CsgpuIndex",
    "const float*": "*const f32", "float*": "*mut f32",
    "const uint32_t*": "*const u32", "uint32_t*": "*mut u32",
    "const uint64_t*": "*const u64", "uint64_t*": "*mut u64",
    "const int32_t*": "*const i32",
    "const csgpu_predicate_t*": "*const CsgpuPredicate", "csgpu_stats_t*": "*mut CsgpuStats",
}


def strip_comments(src: str) -> str:
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def norm_type(t: str) -> str:
    t = re.sub(r"\s+", " ", t.strip())
    t = re.sub(r"\s*\*\s*", "*", t)
    return t


def split_arg(arg: str):
    """'const float *rows' -> ('const float*', 'rows');  'void' -> None;  'csgpu_index *const *peers' -> (..., 'peers')."""
    arg = arg.strip()
    if arg == "void" or not arg:
        return None
    m = re.match(r"^(.*?)([A-Za-z_][A-Za-z_0-9]*)$", arg)
    ctype, name = m.group(1), m.group(2)
    return norm_type(ctype), name


def parse_header(path: str = HEADER):
    src = strip_comments(open(path).read())
    body = src[src.index('extern "C" {'):]
    out = []
    for m in re.finditer(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z_0-9]*\s*\**)\s*(csgpu_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", body, flags=re.M | re.S):
        ret, name, args = norm_type(m.group(1)), m.group(2), m.group(3)
        parsed = [a for a in (split_arg(x) for x in args.replace("\n", " ").split(",")) if a is not None]
        out.append((name, ret, parsed))
    return out


def rust_decl(name, ret, args) -> str:
    rargs = ", ".join(f"{('r#' + a) if a in ('type', 'ref', 'box', 'fn', 'in', 'match') else a}: {CTYPE_TO_RUST[t]}" for t, a in args)
    rret = CTYPE_TO_RUST[ret]
    tail = "" if rret == "()" else f" -> {rret}"
    return f"    pub fn {name}({rargs}){tail};"


def generated_block() -> str:
    lines = [BEGIN]
    for name, ret, args in parse_header():
        lines.append(rust_decl(name, ret, args))
    lines.append(END)
    return "\n".join(lines)


def main():
    src = open(GPU_RS).read()
    a, b = src.index(BEGIN), src.index(END) + len(END)
    new = src[:a] + generated_block() + src[b:]
    if "--check" in sys.argv:
        sys.exit(0 if new == src else 1)
    open(GPU_RS, "w").write(new)
    print(f"{GPU_RS}: {len(parse_header())} declarations")


if __name__ == "__main__":
    main()
