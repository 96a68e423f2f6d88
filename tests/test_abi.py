"""CPU-side checks of the boundary: libtamc.so loads, exports every symbol include/tamc.h declares,
and refuses to compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tamc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tamc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import tamc
    from tamc import binding

    names = _declared()
    assert len(names) >= 20
    L = C.CDLL(tamc.lib_path())
    for n in names:
        assert hasattr(L, n), f"libtamc.so does not export {n}"
    assert sorted(binding.EXPORTS) == names          # the Python binding covers the whole ABI
    assert tamc.lib().tamc_version() == 105


def test_record_and_stats_layout_match_header():
    import tamc

    assert tamc.RECORD_DTYPE.itemsize == 88
    assert C.sizeof(tamc.Stats) == 8 * 10 + 8 * 5 + 8 + 16
    from oracle import oracle as orc

    assert orc.RECORD_DTYPE == tamc.RECORD_DTYPE


def test_no_cpu_fallback():
    import tamc

    if tamc.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(tamc.TamcError) as e:
        tamc.MCTransport(8, 8, 8, 1.0, 1.0, 1.0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_argument_validation_without_device():
    import tamc

    L = tamc.lib()
    h = C.c_void_p()
    assert L.tamc_init(0, 0, 8, 8, 1.0, 1.0, 1.0, 1e-9, C.byref(h)) == 1       # TAMC_EINVAL
    assert b"grid dimensions" in L.tamc_last_error()
    assert L.tamc_init(0, 8, 8, 8, -1.0, 1.0, 1.0, 1e-9, C.byref(h)) == 1
    assert L.tamc_run(None, 10, 1, None, None) == 1
    assert L.tamc_finalize(None) == 0
    assert L.tamc_set_option(None, b"variant", 1) == 1


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (only tests, smoke and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "tissue-ablation-mc_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|oracle/|liboracle|\borc_|tamc_oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(text), (dirpath, f)


def test_every_option_name_is_documented_in_the_header():
    """tamc_set_option / tamc_get_option take names: each one the library knows (option_slot, tamc_api.cu) is described in
    include/tamc.h, the only interface document a binding author reads."""
    import re

    src = open(os.path.join(ROOT, "tissue-ablation-mc_b200", "csrc", "tamc_api.cu")).read()
    body = src[src.index("static int *option_slot("):]
    body = body[:body.index("\n}\n")]
    names = re.findall(r'strcmp\(name, "([a-z0-9_]+)"\)', body)
    assert len(names) >= 30
    header = open(os.path.join(ROOT, "include", "tamc.h")).read()
    missing = [n for n in names if f'"{n}"' not in header]
    assert not missing, missing
