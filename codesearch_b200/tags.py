"""Row tags for the device-side file/language filter (include/csgpu.h csgpu_predicate_t).

Host half of the tag columns: a restatement of the reference's language detection
(`Language::from_path`, /root/reference/src/file/language.rs:31-88; variant order :5-29 gives the
lang_id), its path normalisation (`normalize_path_str`, src/cache/file_meta.rs:23-25) and a dense
file numbering (one file_id per distinct normalised path, in first-seen order — the reference's
FileMetaStore keys files the same way, src/cache/file_meta.rs). A tag is `(lang_id << 27) | file_id`.

The predicate a search carries is built here from what the reference's callers filter on:
  * a path PREFIX (`filter_path`, src/search/mod.rs:727-737, `starts_with` on the root-relative path),
  * a path SUBSTRING (HTTP `path`, src/server/mod.rs:553-559, `contains`),
  * a set of languages (the reference only boosts by `primary_language`, src/search/mod.rs:791-806).
Paths map to a per-FILE bitmap (one bit per file, not per chunk); languages to a 32-bit mask.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np

# src/file/language.rs:5-29, declaration order
LANGUAGES = ("Rust", "Python", "JavaScript", "TypeScript", "Go", "Java", "C", "Cpp", "CSharp", "Ruby", "Php", "Swift",
             "Kotlin", "Shell", "Markdown", "Json", "Yaml", "Toml", "Sql", "Html", "Css", "Xml", "Unknown")
LANG_ID = {name: i for i, name in enumerate(LANGUAGES)}
UNKNOWN = LANG_ID["Unknown"]

TAG_LANG_SHIFT = 27
TAG_FILE_MASK = 0x07FFFFFF
TAG_NONE = 0xFFFFFFFF

# src/file/language.rs:60-88 (extension, lower-cased) and :47-57 (extensionless file names)
_BY_EXT = {}
for _lang, _exts in {
    "Rust": "rs", "Python": "py pyw pyi", "JavaScript": "js mjs cjs", "TypeScript": "ts mts cts tsx jsx", "Go": "go",
    "Java": "java", "C": "c h", "Cpp": "cpp cc cxx hpp hxx", "CSharp": "cs", "Ruby": "rb rake", "Php": "php",
    "Swift": "swift", "Kotlin": "kt kts", "Shell": "sh bash zsh", "Markdown": "md markdown txt", "Json": "json",
    "Yaml": "yaml yml", "Toml": "toml", "Sql": "sql", "Html": "html htm", "Css": "css scss sass less",
    "Xml": "xml csproj props targets resx config",
}.items():
    for _e in _exts.split():
        _BY_EXT[_e] = LANG_ID[_lang]
_BY_NAME = {n: LANG_ID["Shell"] for n in ("Dockerfile", "Containerfile", "Makefile", "GNUmakefile", "makefile", ".env",
                                          ".envrc", "CMakeLists")}
_BY_NAME.update({n: LANG_ID["Ruby"] for n in ("Jenkinsfile", "Vagrantfile", "Fastfile", "Appfile", "Podfile")})


def normalize_path_str(path: str) -> str:
    """src/cache/file_meta.rs:23-25: strip the UNC prefix, backslashes -> forward slashes."""
    while path.startswith("\\\\?\\"):
        path = path[4:]
    return path.replace("\\", "/")


def _rust_extension(file_name: str) -> str:
    """std::path::Path::extension: text after the last '.', none for names like '.env' (leading dot only)."""
    stem = file_name.lstrip(".") if file_name.startswith(".") else file_name
    dot = stem.rfind(".")
    if dot < 0 or file_name in (".", ".."):
        return ""
    return stem[dot + 1:]


def language_from_path(path: str) -> int:
    """lang_id of `Language::from_path` (src/file/language.rs:31-44): extension first, then exact file name."""
    name = normalize_path_str(path).rstrip("/").rsplit("/", 1)[-1]
    ext = _rust_extension(name)
    by_ext = _BY_EXT.get(ext.lower(), UNKNOWN)
    if by_ext != UNKNOWN:
        return by_ext
    return _BY_NAME.get(name, UNKNOWN)


def make_tag(lang_id: int, file_id: int) -> int:
    return ((lang_id & 31) << TAG_LANG_SHIFT) | (file_id & TAG_FILE_MASK)


def synth_tags(first_row: int, n: int) -> np.ndarray:
    """Synthetic tags of csgpu_append_synthetic_tagged (csrc/synth.cuh synth_tags_kernel): file = row // 37,
    lang = fmix32(file) % 23."""
    file = ((np.arange(first_row, first_row + n, dtype=np.uint64) // 37) & TAG_FILE_MASK).astype(np.uint32)
    h = file.copy()
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return ((h % np.uint32(23)) << np.uint32(TAG_LANG_SHIFT)) | file


@dataclass
class TagPredicate:
    """Host form of csgpu_predicate_t. A row passes iff its language bit is set AND file_lo <= file_id <= file_hi
    AND (file_bitmap is None OR bit file_id is set)."""
    lang_mask: int = 0xFFFFFFFF
    file_lo: int = 0
    file_hi: int = 0xFFFFFFFF
    file_bitmap: np.ndarray | None = None   # uint64 words
    n_file_bits: int = 0

    def passes(self, tags: np.ndarray) -> np.ndarray:
        """Reference semantics on the host (used by tests and to build the equivalent id bitmap)."""
        tags = np.asarray(tags, dtype=np.uint32)
        lang = (tags >> np.uint32(TAG_LANG_SHIFT)).astype(np.uint64)
        file = (tags & np.uint32(TAG_FILE_MASK)).astype(np.uint64)
        ok = ((np.uint64(self.lang_mask & 0xFFFFFFFF) >> lang) & np.uint64(1)).astype(bool)
        ok &= (file >= np.uint64(self.file_lo)) & (file <= np.uint64(self.file_hi))
        if self.file_bitmap is not None:
            inb = file < np.uint64(self.n_file_bits)
            words = np.asarray(self.file_bitmap, dtype=np.uint64)
            idx = np.where(inb, file >> np.uint64(6), 0).astype(np.int64)
            bit = (words[idx] >> (file & np.uint64(63))) & np.uint64(1) if words.size else np.zeros_like(file)
            ok &= inb & bit.astype(bool)
        return ok


@dataclass
class FileTable:
    """Dense file numbering: normalised path -> file_id (first-seen order), and its language."""
    ids: dict = field(default_factory=dict)
    paths: list = field(default_factory=list)
    langs: list = field(default_factory=list)

    def file_id(self, path: str) -> int:
        key = normalize_path_str(path)
        fid = self.ids.get(key)
        if fid is None:
            fid = len(self.paths)
            if fid > TAG_FILE_MASK:
                raise ValueError("more than 2^27 files: file_id does not fit the tag")
            self.ids[key] = fid
            self.paths.append(key)
            self.langs.append(language_from_path(key))
        return fid

    @staticmethod
    def from_paths(paths: Sequence[str]) -> "FileTable":
        """Rebuild the table from its persisted path list (file id = position), verbatim."""
        t = FileTable()
        for p in paths:
            key = normalize_path_str(p)
            t.ids.setdefault(key, len(t.paths))   # a duplicate entry keeps its slot so later ids do not shift
            t.paths.append(key)
            t.langs.append(language_from_path(key))
        return t

    def tag(self, path: str) -> int:
        fid = self.file_id(path)
        return make_tag(self.langs[fid], fid)

    def predicate(self, languages: Iterable[str | int] | None = None, path_prefix: str | None = None,
                  path_contains: str | None = None, project_root: str = "") -> TagPredicate:
        """languages: names from LANGUAGES or lang ids. path_prefix follows src/search/mod.rs:727-737 (prefix of the
        root-relative normalised path); path_contains follows src/server/mod.rs:553-559 (substring of the path)."""
        p = TagPredicate()
        if languages is not None:
            p.lang_mask = 0
            for l in languages:
                p.lang_mask |= 1 << (LANG_ID[l] if isinstance(l, str) else int(l))
        if path_prefix is not None or path_contains is not None:
            root = normalize_path_str(project_root)
            flt = normalize_path_str(path_prefix) if path_prefix is not None else None
            mask = np.zeros(len(self.paths), dtype=bool)
            for fid, path in enumerate(self.paths):
                rel = path[len(root):] if root and path.startswith(root) else path
                rel = rel.lstrip("/")
                while rel.startswith("./"):
                    rel = rel[2:]
                ok = True
                if flt is not None:
                    ok = rel.startswith(flt)
                if ok and path_contains is not None:
                    ok = path_contains in path
                mask[fid] = ok
            n = len(self.paths)
            padded = np.zeros(((n + 63) // 64) * 64, dtype=np.uint8)
            padded[:n] = mask
            p.file_bitmap = np.ascontiguousarray(np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint64))
            p.n_file_bits = n
            if n == 0:
                p.file_bitmap = np.zeros(1, dtype=np.uint64)
        return p
