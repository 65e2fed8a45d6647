"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/mwgpu.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from mc_water_ls_mw_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mwgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mwgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 40
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mwgpu.h but not exported by libmwgpu.so"
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding table and header disagree"
    _lib.lib()


def test_struct_sizes_match_header(tmp_path):
    # the header must be plain C, and the ctypes mirrors must have the C compiler's layout
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mwgpu.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mwgpu_mc_params), sizeof(mwgpu_walker_state),'
                   ' offsetof(mwgpu_mc_params, ls), offsetof(mwgpu_walker_state, error), sizeof(mwgpu_flat_params),'
                   ' sizeof(mwgpu_flat_report), offsetof(mwgpu_flat_report, wl_factor));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b, c, d, e, f, g = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert C.sizeof(_lib.McParams) == a and C.sizeof(_lib.WalkerState) == b
    assert _lib.McParams.ls.offset == c and _lib.WalkerState.error.offset == d
    assert C.sizeof(_lib.FlatParams) == e and C.sizeof(_lib.FlatReport) == f and _lib.FlatReport.wl_factor.offset == g


def test_sm100a_code_present():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_fails_loudly_without_gpu():
    L = _lib.lib()
    if L.mwgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = L.mwgpu_create(48, 2, 1, 0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in L.mwgpu_last_error() and b"no CPU fallback" in L.mwgpu_last_error()
    from mc_water_ls_mw_b200 import walkers
    with pytest.raises(_lib.MwgpuError):
        walkers.WalkerBatch(48, 2, 1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may include,
    import, link or dlopen it."""
    bad = re.compile(r"#\s*include[^\n]*oracle|import\s+oracle|from\s+oracle|libmw_oracle|orc\.py|dlopen[^\n]*oracle")
    for top in ("mc_water_ls_mw_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(txt), f"{f} references the oracle"
