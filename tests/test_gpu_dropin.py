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


def _cat_heads(g, pre=""):
    import numpy as np

    return np.ascontiguousarray(np.concatenate([np.concatenate(list(g[f"{pre}{k}"]), axis=1) for k in ("Wq", "Wk", "Wv")], axis=1))


@pytest.mark.parametrize("kind,name", [("enc", "ref_enc_6x8_h4_ff8"), ("enc", "ref_enc_48x64_h4_ff96"),
                                       ("dec", "ref_dec_6x8_h4_ff8_enc6"), ("dec", "ref_dec_40x64_h4_ff96_enc56")])
def test_cpp_encoder_decoder_blocks_match_reference_fixtures(tmp_path, kind, name):
    """The C++ EncoderBlock / DecoderBlock of qg_dropin.cuh (src/transformer.cu:14-168 re-pointed) on the fixture's
    weights: bit-exact against the output of the REFERENCE's kernels; the free functions Encoder() / Decoder() with
    the reference's signatures run and are deterministic for a given seed."""
    import numpy as np

    exe = os.path.join(ROOT, "tests", "cpp", "test_transformer_dropin")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/test_transformer_dropin not built")
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    X = g["X"]
    h, d = X.shape
    E = g["E"] if kind == "dec" else np.zeros((1, d), np.float32)
    if kind == "enc":
        arrays = [X, _cat_heads(g), g["W_O"], g["W1"], g["b1"], g["W2"], g["b2"]]
    else:
        arrays = [X, E, _cat_heads(g, "sa_"), g["sa_W_O"], _cat_heads(g, "ca_"), g["ca_W_O"], g["W1"], g["b1"], g["W2"], g["b2"]]
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(src, "wb") as f:
        f.write(np.array([h, E.shape[0], d, int(g["heads"]), int(g["d_ff"])], np.int32).tobytes())
        for a in arrays:
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    r = subprocess.run([exe, kind, str(src), str(dst)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(dst, "rb").read()
    out = np.frombuffer(raw[: 4 * h * d], np.float32).reshape(h, d)
    ok = int(np.frombuffer(raw[4 * h * d:], np.int32)[0])
    exp = g["out"]
    nan = np.isnan(exp)
    assert np.array_equal(nan, np.isnan(out)) and np.array_equal(out.view(np.int32)[~nan], exp.view(np.int32)[~nan])
    assert ok == 1


@pytest.mark.parametrize("flags", [[], ["-m", "256", "-n", "128", "-k", "192", "-s", "5"], ["-m", "300", "-n", "200", "-k", "100"]])
def test_timing_driver_op_by_op_on_the_new_path(flags):
    """src/timing_quantize.cu (the benchmark of record; its committed merge conflict resolved to the README's
    2048 x 512 x 512) stock, and with EVERY op of its inlined quantized sequence re-pointed one by one through
    qg_dropin.cuh (op_mm fp32 / int8 / K=1 outer product, op_absmax both directions, op_inv_divide, quantizing
    op_multiply, op_dequantize, op_multiply by a constant, op_subtract): the mean quantization error it prints for
    each of its 50 iterations must be the same text."""
    def errors(path):
        if not os.path.exists(path):
            pytest.skip(f"{os.path.relpath(path, ROOT)} not built")
        r = subprocess.run([path, *flags], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        lines = r.stdout.splitlines()
        return [lines[i + 1] for i, l in enumerate(lines) if l.startswith("Mean Quantization error")]

    ref = errors(os.path.join(ROOT, "oracle", "_ref", "timing_quantize_ref"))
    new = errors(os.path.join(ROOT, "oracle", "_ref", "timing_quantize_dropin"))
    assert len(ref) == 50 and new == ref
