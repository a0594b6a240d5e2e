"""Drop-in check at the reference's own call site: src/test_quantize.cu compiled unmodified
(oracle/_ref/test_quantize_ref) and with op_quantized_mm's body re-pointed at libqgemm.so through
qg_dropin.cuh (oracle/_ref/test_quantize_dropin; see oracle/make_dropin_demo.py and INTEGRATION.md)
must print the same results.  Also runs the C++ driver written against the drop-in layer."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} not built")
    r = subprocess.run([path], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_reference_driver_prints_identical_results_on_the_new_path():
    ref = run(os.path.join(ROOT, "oracle", "_ref", "test_quantize_ref"))
    new = run(os.path.join(ROOT, "oracle", "_ref", "test_quantize_dropin"))
    assert "Quantized result" in ref and "All tests completed successfully!" in ref
    assert new == ref


def test_reference_softmax_driver_prints_identical_results_on_the_new_path():
    ref = run(os.path.join(ROOT, "oracle", "_ref", "test_softmax_ref"))
    new = run(os.path.join(ROOT, "oracle", "_ref", "test_softmax_dropin"))
    assert "Test passed." in ref
    assert new == ref


def test_cpp_driver_over_dropin_layer():
    out = run(os.path.join(ROOT, "tests", "cpp", "test_quantize_dropin"))
    assert "All tests completed successfully!" in out
