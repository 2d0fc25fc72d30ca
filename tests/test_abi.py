"""CPU suite, part 2: the C-ABI library loads, exports every symbol that
include/b200grav.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "b200grav.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    import b200grav
    lib = b200grav.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200grav.h but not exported"
    assert sorted(b200grav.EXPORTS) == names
    assert lib.b200_abi_version() == 2


def test_host_scalars_match_oracle(oracle):
    import b200grav
    lib = b200grav.load_library()
    for a in (0.02, 0.3, 1.0, 1.9):
        assert lib.b200_hubble_a(a, 0.31, 0.0, 0.69, 0.67) == oracle.hubble_a(a)
        assert lib.b200_scale_factor_step(a, 1e-3, 0.31, 0.0, 0.69, 0.67) == oracle.scale_factor_step(a, 1e-3)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import b200grav
    with pytest.raises(b200grav.B200Error, match="no usable sm_100"):
        b200grav.Engine(0)
    lib = b200grav.load_library()
    pos = np.zeros((4, 3), np.float32)
    out = np.zeros((4, 3), np.float32)
    assert lib.b200_direct_forces_host(None, pos.ctypes.data, None, out.ctypes.data, 4, 0.01, 0.0) == 1


def test_missing_library_raises(tmp_path):
    import b200grav
    with pytest.raises(b200grav.B200Error, match="no CPU fallback"):
        b200grav.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    """The product tree must not reference oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "lambda-cdm-raytracing_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for tok in ("pyoracle", "liboracle", "oracle/", "oracle.h", "orc_", "lcdm_ref"):
                    assert tok not in txt, (os.path.join(dp, f), tok)


def test_shard_range_matches_python_and_partitions():
    import b200grav
    lib = b200grav.load_library()
    for n in (0, 1, 7, 12345, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for rank in range(world):
                i0, nl = C.c_size_t(), C.c_size_t()
                assert lib.b200_shard_range(n, rank, world, C.byref(i0), C.byref(nl)) == 0
                lo, hi = b200grav.shard_range(n, rank, world)
                assert (i0.value, nl.value) == (lo, hi - lo)
                assert i0.value == covered
                covered += nl.value
            assert covered == n
    assert lib.b200_shard_range(10, 2, 2, None, None) == 1      # rank out of range


def test_ic_params_layout_matches_header(tmp_path):
    """The ctypes mirror of b200_ic_params has the header's size and field offsets (checked with gcc, no GPU), and
    b200_ic_params_default -- a pure host function -- fills the reference's defaults (initial_conditions.hpp:19-45)."""
    import ctypes as C
    import subprocess
    import b200grav
    fields = [f[0] for f in b200grav.ICParams._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200grav.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(b200_ic_params));\n' +
                   "".join(f'  printf("%zu\\n", offsetof(b200_ic_params, {f}));\n' for f in fields) +
                   "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == C.sizeof(b200grav.ICParams)
    assert out[1:] == [getattr(b200grav.ICParams, f).offset for f in fields]
    lib = b200grav.load_library()
    p = b200grav.ICParams()
    lib.b200_ic_params_default(C.byref(p))
    assert (p.box, p.z_initial, p.seed, p.use_2lpt) == (100.0, 49.0, 12345, 0)
    assert (p.omega_m, p.omega_lambda, p.omega_k, p.h, p.sigma_8, p.n_s) == (0.31, 0.69, 0.0, 0.67, 0.81, 0.965)


def test_cpp_example_fails_loudly_without_gpu():
    """The C++ product path (B200LambdaCDMSimulation behind examples/nbody_b200) has no CPU fallback either: on a box
    without a GPU it must exit non-zero with a clear message, not produce numbers."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "lambda-cdm-raytracing_b200", "examples", "_bin", "nbody_b200")
    if torch.cuda.is_available() or not os.path.exists(exe):
        pytest.skip("needs a GPU-less box and the built example")
    r = subprocess.run([exe, "1000", "2", "tree"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
