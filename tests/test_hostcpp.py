"""The C++ host layer (centrolign_b200/hostcpp/po_poa_b200.hpp) that keeps the reference's po_poa signature.

CPU: it compiles with plain g++ -std=c++11 (the reference's dialect, CMakeLists.txt:9) and links.
GPU: the reference's unit-test goldens through the wrapper, and -- when the binary built against the
reference's own headers travelled with the snapshot -- BaseGraph / AlignmentParameters / Alignment
straight through the wrapper, compared with centrolign::po_poa."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_hostcpp")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "test_dropin")


def _build():
    lib = os.path.join(ROOT, "centrolign_b200", "csrc", "libcentrolign_b200.so")
    if not os.path.exists(lib):
        pytest.skip("libcentrolign_b200.so not built")
    src = os.path.join(ROOT, "tests", "cpp", "test_hostcpp.cpp")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.run(["g++", "-std=c++11", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "centrolign_b200", "hostcpp"), src,
                        "-L" + os.path.dirname(lib), "-lcentrolign_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", BIN],
                       check=True)


def test_wrapper_compiles_and_links():
    _build()
    out = subprocess.run([BIN, "--link-only"], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert "linked" in out


def test_fill_in_pool_rendezvous():
    """hostcpp/chain_batcher.hpp without a device: every job of a pool of worker threads runs exactly once, its worker
    continues only after it ran, launches are shared, a pool of one is the serial loop, and an error in one launch comes
    back as an exception."""
    _build()
    res = subprocess.run([BIN, "--batcher"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert res.returncode == 0 and "batcher passed all tests!" in res.stdout, res.stdout


@pytest.mark.gpu
def test_wrapper_reference_goldens():
    _build()
    res = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0 and "passed all tests!" in res.stdout, res.stdout


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/test_dropin did not travel")
def test_dropin_with_reference_types():
    res = subprocess.run([DROPIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0 and "passed all tests!" in res.stdout, res.stdout
