"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
include/sgb200.h declares; no compute call is made (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from speakerguard_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from speakerguard_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in sgb200.h but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == syms


def test_version_and_pure_host_helpers(lib):
    assert lib.sg_version() == 200
    # kaldi.py:70: m = (N + 80) // 160
    for n, m in [(32000, 200), (48000, 300), (80000, 500), (17777, 111), (400, 3)]:
        assert lib.sg_num_frames(n) == m


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from speakerguard_b200 import _lib
    h = ctypes.c_void_p()
    rc = lib.sg_create(ctypes.byref(h), 0)
    assert rc == _lib.SG_ECUDA
    assert b"no CPU fallback" in lib.sg_last_error()
    from speakerguard_b200.engine import Engine
    with pytest.raises(_lib.SgError):
        Engine("cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "speakerguard_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("sg_oracle-free", ""), f"{f} mentions the oracle"


def test_python_constants_mirror_the_header():
    """The ctypes layer restates a few enums of include/sgb200.h; they must not drift."""
    from speakerguard_b200 import _lib
    src = open(os.path.join(ROOT, "include", "sgb200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(SG_OPT_[A-Z0-9_]+)\s+(\d+)", src)}
    assert defs == {"SG_OPT_POOL_FUSION": _lib.OPT_POOL_FUSION, "SG_OPT_FEAT_STASH": _lib.OPT_FEAT_STASH,
                    "SG_OPT_L1_TAP_FORM": _lib.OPT_L1_TAP_FORM, "SG_OPT_UTT_OFFSET": _lib.OPT_UTT_OFFSET,
                    "SG_OPT_CUDA_GRAPH": _lib.OPT_CUDA_GRAPH, "SG_OPT_CMVN_FUSION": _lib.OPT_CMVN_FUSION, "SG_OPT_ROW_COMPACTION": _lib.OPT_ROW_COMPACTION}
    body = re.search(r"enum\s*\{\s*SG_PROF_MFCC_FWD\s*=\s*0(.*?)SG_PROF_COUNT\s*\}", re.sub(r"/\*.*?\*/", "", src, flags=re.S), re.S)
    assert body is not None
    n_prof = 1 + len(re.findall(r"SG_PROF_[A-Z0-9_]+", body.group(1)))
    assert n_prof == _lib.PROF_COUNT
