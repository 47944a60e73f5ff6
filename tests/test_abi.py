"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/reyna_b200.h declares
(no compute calls: those need a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "reyna_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(reyna_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for n in ("reyna_b200_structure", "reyna_b200_csr_pattern", "reyna_b200_assemble", "reyna_b200_boundary",
              "reyna_b200_compact_count", "reyna_b200_compact_fill", "reyna_b200_solve", "reyna_b200_spmv",
              "reyna_b200_errors", "reyna_b200_last_error"):
        assert n in names


def test_library_builds_loads_and_exports_every_declared_symbol():
    from reyna_b200 import _lib
    _lib.build()
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/reyna_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared_functions()


def test_ctypes_structs_match_header_sizes():
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    import subprocess
    import tempfile
    from reyna_b200 import _lib
    names = [("rb_geom", _lib.RbGeom), ("rb_quad", _lib.RbQuad), ("rb_coefs", _lib.RbCoefs),
             ("rb_structure", _lib.RbStructure), ("rb_solve_opts", _lib.RbSolveOpts), ("rb_solve_info", _lib.RbSolveInfo),
             ("rb_err_samples", _lib.RbErrSamples), ("rb_mg_opts", _lib.RbMgOpts), ("rb_mg_dist", _lib.RbMgDist)]
    names = [(n, t) for n, t in names if re.search(r"\b%s\b" % n, open(HEADER).read())]
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        with open(src, "w") as fh:
            fh.write('#include <stdio.h>\n#include "reyna_b200.h"\nint main(void){\n')
            for n, _ in names:
                fh.write(f'printf("%zu\\n", sizeof({n}));\n')
            fh.write("return 0;}\n")
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    for (n, t), size in zip(names, out):
        assert ctypes.sizeof(t) == int(size), n


def test_ctypes_field_offsets_match_header():
    """offsetof() of every field of the option / distribution structs (a reordered field keeps the size)."""
    import subprocess
    import tempfile
    from reyna_b200 import _lib
    structs = [("rb_mg_opts", _lib.RbMgOpts), ("rb_mg_dist", _lib.RbMgDist), ("rb_solve_opts", _lib.RbSolveOpts),
               ("rb_solve_info", _lib.RbSolveInfo), ("rb_structure", _lib.RbStructure), ("rb_coefs", _lib.RbCoefs),
               ("rb_geom", _lib.RbGeom)]
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "o.c")
        with open(src, "w") as fh:
            fh.write('#include <stdio.h>\n#include <stddef.h>\n#include "reyna_b200.h"\nint main(void){\n')
            for n, t in structs:
                for f in t._fields_:
                    fh.write(f'printf("%zu\\n", offsetof({n}, {f[0]}));\n')
            fh.write("return 0;}\n")
        exe = os.path.join(td, "o")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = iter(subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    for n, t in structs:
        for f in t._fields_:
            assert getattr(t, f[0]).offset == int(next(out)), (n, f[0])


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import reyna_b200 as rb
    geo = rb.DGFEMGeometry(rb.voronoi_mesh(32, seed=1))
    dg = rb.DGFEM(geo, polynomial_degree=1)
    dg.add_data(advection=lambda x: np.ones(x.shape, dtype=np.float64),
                dirichlet_bcs=lambda x: np.ones(x.shape[0], dtype=np.float64))
    with pytest.raises(rb.ReynaB200Error):
        dg.dgfem(solve=True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reyna_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn
