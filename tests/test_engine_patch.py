"""CPU check of the N1 artefact: integration/engine_wiring.patch applies cleanly to the reference's own
engine sources and the patched translation unit compiles against the reference's headers and this repository's
plugin headers.  Needs /root/reference (absent on the GPU box: skipped there; the executable test of the patched
engine is tests/test_gpu_host_plugin.py::test_engine_wiring_builder_run)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "core")), reason="reference sources not present")
def test_engine_wiring_patch_applies_and_compiles():
    tmp = tempfile.mkdtemp(prefix="b200_engine_patch_")
    try:
        os.makedirs(os.path.join(tmp, "include", "core"))
        os.makedirs(os.path.join(tmp, "src", "core"))
        shutil.copy(os.path.join(REF, "include", "core", "simulation_engine.hpp"), os.path.join(tmp, "include", "core"))
        shutil.copy(os.path.join(REF, "src", "core", "simulation_engine.cpp"), os.path.join(tmp, "src", "core"))
        patch = os.path.join(ROOT, "integration", "engine_wiring.patch")
        r = subprocess.run(["patch", "-s", "-p1", "-i", patch], cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert not [f for _, _, fs in os.walk(tmp) for f in fs if f.endswith((".rej", ".orig"))]
        src = open(os.path.join(tmp, "src", "core", "simulation_engine.cpp")).read()
        for hook in ("void SimulationEngine::compute_forces", "void SimulationEngine::integrate_step",
                     "void SimulationEngine::update_cosmology"):
            body = src[src.index(hook):]
            body = body[body.index("{") + 1:body.index("\n}")]
            assert len(body.strip()) > 0, hook + " is still empty"
        inc = ["-I" + os.path.join(tmp, "include"), "-I" + os.path.join(REF, "include"),
               "-I" + os.path.join(REF, "include", "core"), "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"),
               "-I" + os.path.join(ROOT, "lambda-cdm-raytracing_b200", "host")]
        for defs in ([], ["-DHAVE_B200GRAV"]):
            r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w"] + defs + inc +
                               [os.path.join(tmp, "src", "core", "simulation_engine.cpp")], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
