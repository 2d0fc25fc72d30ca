"""-m gpu: the C++ plugin layer (IForceComputer / IIntegrator / ICosmologyModel
adapters + ForceComputerFactory registration) exercised from a C++ program that
also links the reference's CPU TreeForceComputer as the checker."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_host_plugin_parity():
    exe = os.path.join(ROOT, "tests", "host", "_bin", "host_parity_test")
    if not os.path.exists(exe):
        pytest.skip("tests/host/_bin/host_parity_test not built (needs the reference headers at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:]
    assert "HOST PARITY OK" in r.stdout and "FAIL" not in r.stdout


def test_host_sharded_simulation_matches_single_gpu():
    """C++ only, no Python in the loop: one B200LambdaCDMSimulation per GPU (threads), NCCL
    all-gather through the C ABI, against the 1-GPU run.  Skips itself on a 1-GPU box."""
    exe = os.path.join(ROOT, "tests", "host", "_bin", "shard_test")
    if not os.path.exists(exe):
        pytest.skip("tests/host/_bin/shard_test not built (needs the reference headers at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout
    if "SKIP" in r.stdout:
        pytest.skip(r.stdout.strip())


@pytest.mark.parametrize("args", [("20000", "20", "direct", "random"), ("32768", "20", "tree-fixed", "zeldovich"),
                                  ("20000", "10", "tree", "random"), ("20000", "10", "tree-periodic", "random")])
def test_example_program(args):
    """The reference's cuda_nbody_test flow (examples/cuda_nbody_test.cpp) on the B200 simulation class."""
    exe = os.path.join(ROOT, "lambda-cdm-raytracing_b200", "examples", "_bin", "nbody_b200")
    if not os.path.exists(exe):
        pytest.skip("examples/_bin/nbody_b200 not built (needs the reference headers at build time)")
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-1000:])
    assert r.returncode == 0, r.stderr[-1000:]
    assert "particle-updates/second" in r.stdout and "nan" not in r.stdout.lower()


def test_engine_wiring_builder_run():
    """the reference's SimulationBuilder / SimulationEngine with integration/engine_wiring.patch applied
    (to a temporary copy of its sources, at build time) runs 10 KDK steps through DirectForceComputer /
    TreeForceComputer on the B200 -- device-resident and through host arrays -- and lands on the CPU reference's
    positions."""
    exe = os.path.join(ROOT, "tests", "host", "_bin", "engine_wiring_test")
    if not os.path.exists(exe):
        pytest.skip("tests/host/_bin/engine_wiring_test not built (needs the reference sources at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ENGINE WIRING OK" in r.stdout and "FAIL" not in r.stdout
