"""The C-ABI boundary without a GPU: the library builds for sm_100a, loads, exports every symbol that
include/som_b200.h declares, the ctypes signatures mirror the header argument for argument, and the entry points
fail loudly (no fallback) when there is no B200."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "som_b200.h")


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def declared_prototypes():
    """{name: [C type of every parameter]} parsed from the header (comments stripped, parameter names dropped)."""
    src = _strip_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|void|const char\*)\s+(som_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        types = []
        if params not in ("", "void"):
            for prm in params.split(","):
                prm = " ".join(prm.split())
                t = re.sub(r"\s*\b\w+$", "", prm) if not prm.endswith("*") else prm      # drop the parameter name
                types.append(t.replace(" *", "*"))
        out[name] = types
    return out


def declared_functions():
    """{name: number of parameters}."""
    return {k: len(v) for k, v in declared_prototypes().items()}


def ctype_of(c_type):
    """The ctypes type a C parameter type must be bound with (every pointer travels as void*)."""
    if c_type.endswith("*"):
        return ctypes.c_void_p
    return {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
            "unsigned int": ctypes.c_uint32}[c_type]


@pytest.fixture(scope="module")
def lib():
    from vit_som_b200 import _lib
    _lib.build()                       # no-op when libsom_b200.so is newer than its sources
    return _lib.lib()


def test_header_symbols_are_exported_and_bound(lib):
    from vit_som_b200 import _lib
    decl = declared_functions()
    assert len(decl) >= 25 and "som_forward" in decl and "som_backward_dx" in decl
    for name, nparams in decl.items():
        assert hasattr(lib, name), f"{name} is declared in som_b200.h but not exported by libsom_b200.so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in vit_som_b200/_lib.py"
        assert len(_lib.SIGNATURES[name][1]) == nparams, f"{name}: header has {nparams} parameters"
    for name in _lib.SIGNATURES:
        assert name in decl, f"{name} is bound in _lib.py but not declared in include/som_b200.h"


def test_ctypes_signatures_match_the_header_type_for_type():
    """Every argument of every binding has the ctypes type of the header's parameter at that position (an int where
    the header has an int64_t - or two swapped arguments of different types - would corrupt the call silently)."""
    from vit_som_b200 import _lib
    for name, c_types in declared_prototypes().items():
        bound = _lib.SIGNATURES[name][1]
        for i, (c_t, b) in enumerate(zip(c_types, bound)):
            assert ctype_of(c_t) is b, f"{name} argument {i}: header `{c_t}` but bound as {b.__name__}"


def _integration_md():
    return open(os.path.join(ROOT, "INTEGRATION.md")).read()


def test_integration_doc_ctypes_stub_matches_the_header():
    """The ctypes stub printed in INTEGRATION.md declares som_forward with the header's parameter list (the round-1
    document had lost two parameters: a maintainer who followed it passed the stream where the workspace goes)."""
    md = _integration_md()
    m = re.search(r"L\.som_forward\.argtypes\s*=\s*\[(.*?)\]", md, flags=re.S)
    assert m, "INTEGRATION.md no longer shows the som_forward ctypes stub"
    body = re.sub(r"#[^\n]*", "", m.group(1))
    names = [t.strip() for t in body.replace("\n", " ").split(",") if t.strip()]
    alias = {"P": ctypes.c_void_p, "I64": ctypes.c_int64, "I": ctypes.c_int, "F": ctypes.c_float}
    doc_types = [alias[n] for n in names]
    header_types = [ctype_of(t) for t in declared_prototypes()["som_forward"]]
    assert doc_types == header_types
    # the call in the stub passes as many arguments as the header has parameters
    call = re.search(r"rc = L\.som_forward\((.*?)\)\n\s*if rc", md, flags=re.S)
    assert call
    depth, nargs = 0, 1
    for ch in call.group(1):
        depth += ch in "([" 
        depth -= ch in ")]"
        nargs += ch == "," and depth == 0
    assert nargs == len(header_types)


def test_integration_doc_cpp_stub_matches_the_header():
    md = _integration_md()
    call = re.search(r"int rc = som_forward\((.*?)\);\n", md, flags=re.S)
    assert call, "INTEGRATION.md no longer shows the libtorch stub"
    depth, nargs = 0, 1
    for ch in call.group(1):
        depth += ch in "(<"
        depth -= ch in ")>"
        nargs += ch == "," and depth == 0
    assert nargs == len(declared_prototypes()["som_forward"])


def test_abi_version_and_pure_host_entry_points(lib):
    assert lib.som_b200_abi_version() == 7
    assert lib.som_gemm_workspace_floats() == 4096 + 2 * 74 * 256 * 256  # 74 CTA pairs (a B200, and the GPU-less default), two slots each
    assert lib.som_loss_scratch_floats(1024, 1600) >= 1024 * 2
    assert lib.som_loss_fused_scratch_floats(1024, 1600) == (1024 // 8) * 4 + 2
    lib.som_launch_count_reset()
    assert lib.som_launch_count() == 0


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less build container")
    buf = (ctypes.c_float * 16)()
    rc = lib.som_prep_rows(ctypes.addressof(buf), 1, 4, 4, 0, ctypes.addressof(buf), ctypes.addressof(buf), 4,
                           ctypes.addressof(buf), None)
    assert rc != 0 and lib.som_last_error()
    from vit_som_b200 import SOMLayer, SomError
    from oracle.ref_import import make_config
    layer = SOMLayer(make_config([4, 4], 8))
    with pytest.raises(SomError):
        layer(torch.randn(3, 8))


def _bounds(lib, tiles0, nkb0, tiles1, nkb1, workers, split):
    import ctypes
    out = (ctypes.c_int64 * (workers + 1))()
    assert lib.som_debug_schedule(tiles0, nkb0, tiles1, nkb1, workers, split, out) == 0
    return list(out)


def test_streamk_ranges_partition_the_work(lib):
    """The k-block ranges of the CTA pairs (host view of the device scheduler) cover every unit exactly once, in
    order, for one and two GEMMs per launch."""
    import random
    rng = random.Random(0)
    for _ in range(200):
        tiles0, nkb0 = rng.randint(1, 300), rng.randint(1, 200)
        two = rng.random() < 0.5
        tiles1, nkb1 = (rng.randint(1, 300), rng.randint(1, 200)) if two else (0, 0)
        workers = rng.randint(2, 74)
        b = _bounds(lib, tiles0, nkb0, tiles1, nkb1, workers, 0)
        total = tiles0 * nkb0 + tiles1 * nkb1
        assert b[0] == 0 and b[-1] == total
        assert all(x <= y for x, y in zip(b, b[1:]))
        sizes = [y - x for x, y in zip(b, b[1:])]
        assert max(sizes) - min(sizes) <= 1                       # even ranges


def test_splitk_pieces_are_tile_aligned(lib):
    """Split-K: worker p = tile * split + piece; pieces of a tile are contiguous, the owner's (piece 0) starts at
    the tile's first k-block and is the longest (it waits for the others)."""
    for tiles, nkb, split in [(36, 98, 2), (10, 40, 3), (5, 64, 2), (7, 200, 4)]:
        workers = tiles * split
        b = _bounds(lib, tiles, nkb, 0, 0, workers, split)
        assert b[0] == 0 and b[-1] == tiles * nkb
        assert all(x <= y for x, y in zip(b, b[1:]))
        for t in range(tiles):
            assert b[t * split] == t * nkb                        # owner starts at the tile's head
            pieces = [b[t * split + i + 1] - b[t * split + i] for i in range(split)]
            assert sum(pieces) == nkb and pieces[0] == max(pieces)


def test_every_worker_of_a_cut_tile_has_work(lib):
    """Hand-over invariant of the in-kernel reduction: the owner of a cut tile waits for the flag of EVERY worker whose
    range begins inside the tile, so no such range may be empty (the library only builds schedules with at least 4
    k-blocks per stream-K worker and at least 8 per split-K piece)."""
    import random
    rng = random.Random(1)
    for _ in range(200):
        tiles, nkb = rng.randint(1, 200), rng.randint(8, 300)
        workers = rng.randint(2, min(74, max(2, tiles * nkb // 4)))
        b = _bounds(lib, tiles, nkb, 0, 0, workers, 0)
        assert all(y - x >= 4 or tiles * nkb < 4 * workers for x, y in zip(b, b[1:]))
    for tiles, nkb in [(36, 98), (20, 64), (9, 33), (74, 16), (3, 1000)]:
        split = min(74 // tiles, nkb // 8)
        if split < 2:
            continue
        b = _bounds(lib, tiles, nkb, 0, 0, tiles * split, split)
        assert all(y - x >= 1 for x, y in zip(b, b[1:]))


def test_two_phase_schedule_partitions_both_gemms(lib):
    """Data-parallel backward: every worker gets an even share of GEMM 0 (dW) and then an even share of GEMM 1 (dx);
    each phase covers its GEMM exactly once, in order, and no share is empty when there are at least as many units as
    workers (the hand-over invariant of the in-kernel reduction, per phase)."""
    import random
    rng = random.Random(2)
    for _ in range(200):
        tiles0, nkb0 = rng.randint(1, 300), rng.randint(1, 200)
        tiles1, nkb1 = rng.randint(1, 300), rng.randint(1, 200)
        workers = rng.randint(2, 74)
        out = (ctypes.c_int64 * (2 * (workers + 1)))()
        assert lib.som_debug_schedule(tiles0, nkb0, tiles1, nkb1, workers, -1, out) == 0
        b = list(out)
        ph0, ph1 = b[:workers + 1], b[workers + 1:]
        u0, u1 = tiles0 * nkb0, tiles1 * nkb1
        assert ph0[0] == 0 and ph0[-1] == u0 and ph1[0] == u0 and ph1[-1] == u0 + u1
        for ph, units in ((ph0, u0), (ph1, u1)):
            sizes = [y - x for x, y in zip(ph, ph[1:])]
            assert all(sz >= 0 for sz in sizes) and max(sizes) - min(sizes) <= 1
            assert units < workers or min(sizes) >= 1
    out = (ctypes.c_int64 * 8)()
    assert lib.som_debug_schedule(4, 8, 0, 0, 3, -1, out) != 0          # needs two GEMMs
