"""CPU tests of the drop-in boundary: the library loads, exports every symbol include/mceik_b200.h
declares, the struct layouts are those of the reference header, and compute fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mceik_b200.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:mceik_|eikonal3d_|locate3d_|locate_|computeHomog|weightedMedian__)\w*)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mceik_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mceik_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_header_compiles_as_c_and_layouts_match_reference(tmp_path):
    """The header is plain C; struct offsets equal those of an LP64 build of the reference's mceik_struct.h."""
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "mceik_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(struct mceik_catalog_struct), offsetof(struct mceik_catalog_struct, tobs),
         offsetof(struct mceik_catalog_struct, obsPtr), offsetof(struct mceik_catalog_struct, nevents),
         sizeof(struct mceik_stations_struct));
  printf("%zu %zu %zu %zu\n", offsetof(struct mceik_stations_struct, xrec), offsetof(struct mceik_stations_struct, nstat),
         sizeof(struct mceik_parms_struct), offsetof(struct mceik_parms_struct, x0));
  printf("%zu %zu\n", sizeof(mceik_fsm_grid), offsetof(mceik_fsm_grid, maxit));
  return 0; }''')
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert out[:5] == ["96", "32", "80", "88", "96"]
    assert out[5:9] == ["32", "88", "1256", "1184"]
    ref_hdr = "/root/reference/include/mceik_struct.h"
    if os.path.exists(ref_hdr):  # same numbers from the reference's own header, when it is around
        src2 = tmp_path / "r.c"
        src2.write_text(src.read_text().replace('#include "mceik_b200.h"', f'#include "{ref_hdr}"\n#include "mceik_b200.h"'))
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src2), "-o", str(exe)])
        assert subprocess.check_output([str(exe)], text=True).split() == out
    from mceik_b200 import _lib
    assert C.sizeof(_lib.CatalogStruct) == 96 and C.sizeof(_lib.StationsStruct) == 96
    assert C.sizeof(_lib.FsmGrid) == int(out[9]) and _lib.FsmGrid.maxit.offset == int(out[10])


def test_no_cpu_fallback():
    """Without a usable B200 every entry point reports an error; nothing is computed on the CPU."""
    import numpy as np
    import mceik_b200
    from mceik_b200 import _lib
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.MceikError, match="no CUDA device"):
        mceik_b200.Context(0)
    from mceik_b200 import eikonal as E
    u = np.zeros(8)
    assert E.eikonal3d_serial_driver(1, 0, 1, 1, 2, 2, 2, 1e-6, 1.0, 0, 0, 0, [0.0], [0.5], [0.5], [0.5], np.ones(8), u) == 1
    from mceik_b200 import locate as L
    x = np.arange(4.0)
    t0, obj = L.aligned_empty(64, np.float64), L.aligned_empty(64, np.float64)
    test = L.aligned_empty(128, np.float64)
    assert L.locate_l2_gridSearch__double64(64, 64, 2, 1, 0.0, np.zeros(2, np.int32), x[:2], None, np.ones(2), test, t0, obj) == 1
    assert L.locate3d_initialize() == 1


def test_product_does_not_import_oracle():
    """Nothing under mceik_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mceik_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)


def test_c_host_example_compiles_and_links(tmp_path):
    """tests/c_host_example.c: a plain C driver against include/mceik_b200.h links to the library with gcc; without a
    B200 it fails loudly (no CPU fallback) instead of computing anything."""
    import subprocess
    from test_gpu_c_host import build_c_host
    exe = build_c_host(tmp_path)
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode == 1 and "no CPU fallback" in out.stdout
