"""Host half of the row-tag filter: language detection and predicate construction
against the oracle's independent restatement and the reference's own known answers
(/root/reference/src/file/language.rs:145-166). CPU only."""
import numpy as np

from codesearch_b200 import tags as T


def test_language_known_answers(oracle):
    # the reference's tests: main.rs -> Rust; py/pyi -> Python; ts/tsx/jsx -> TypeScript (language.rs:145-166)
    kat = {"main.rs": "Rust", "a.py": "Python", "a.pyi": "Python", "x.ts": "TypeScript", "x.tsx": "TypeScript",
           "x.jsx": "TypeScript"}
    for path, lang in kat.items():
        assert oracle.language_from_path(path) == lang
        assert T.LANGUAGES[T.language_from_path(path)] == lang


def test_language_matches_oracle(oracle):
    paths = ["src/lib.rs", "a/b/c.PY", "Dockerfile", "x/Makefile", ".env", ".envrc", "CMakeLists.txt", "CMakeLists",
             "noext", "dir.d/file", "a.tar.gz", "k.kts", "web/index.HTML", "s.scss", "proj.csproj", "app.config",
             "C:\\repo\\src\\main.cpp", "\\\\?\\C:\\x\\y.go", ".hidden.json", "Podfile", "README.md", "notes.txt", "q.sql",
             "trailing/", "a.b.c.yml", ".gitignore", "x.", "weird.RS"]
    assert T.LANGUAGES == tuple(oracle.LANGUAGE_ORDER)
    for p in paths:
        assert T.LANGUAGES[T.language_from_path(p)] == oracle.language_from_path(p), p


def test_synth_tags_match_oracle(oracle):
    for first in (0, 35, 1 << 20, (1 << 32) + 5):
        assert np.array_equal(T.synth_tags(first, 200), oracle.synth_tags(first, 200))


def test_predicate_matches_oracle(oracle):
    rng = np.random.default_rng(5)
    tags = oracle.synth_tags(0, 4000)
    tags[::97] = T.TAG_NONE
    n_files = 4000 // 37 + 1
    fmask = rng.random(n_files) < 0.3
    bm = np.zeros((n_files + 63) // 64, dtype=np.uint64)
    for f in np.nonzero(fmask)[0]:
        bm[f >> 6] |= np.uint64(1) << np.uint64(f & 63)
    cases = [
        dict(),
        dict(lang_mask=(1 << 0) | (1 << 5) | (1 << 22)),
        dict(file_lo=10, file_hi=57),
        dict(lang_mask=0x7FFFFF, file_bitmap=bm, n_file_bits=n_files),
        dict(file_bitmap=bm, n_file_bits=50),
        dict(lang_mask=0),
    ]
    for c in cases:
        want = oracle.tag_predicate_mask(tags, **c)
        got = T.TagPredicate(**c).passes(tags)
        assert np.array_equal(want, got), c


def test_file_table_predicate():
    ft = T.FileTable()
    paths = ["/repo/src/a.rs", "/repo/src/b.py", "/repo/docs/x.md", "/repo/src/sub/c.rs", "/repo/Makefile"]
    tags = np.array([ft.tag(p) for p in paths], dtype=np.uint32)
    assert [t & T.TAG_FILE_MASK for t in tags.tolist()] == [0, 1, 2, 3, 4]
    p = ft.predicate(languages=["Rust"])
    assert p.passes(tags).tolist() == [True, False, False, True, False]
    p = ft.predicate(path_prefix="src/", project_root="/repo")           # src/search/mod.rs:727-737
    assert p.passes(tags).tolist() == [True, True, False, True, False]
    p = ft.predicate(path_contains="sub")                                 # src/server/mod.rs:553-559
    assert p.passes(tags).tolist() == [False, False, False, True, False]
    p = ft.predicate(languages=["Rust", "Shell"], path_prefix="src", project_root="/repo/")
    assert p.passes(tags).tolist() == [True, False, False, True, False]
    assert ft.tag("/repo/src/a.rs") == tags[0]                            # stable numbering
