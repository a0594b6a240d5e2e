"""Encoder / Decoder of src/transformer.cu:14-168 on the quantized path, weights persistent.

The reference functions re-draw every weight with op_uniform_init on every call, loop the heads with
one AttentionLayer each, concatenate the heads through the host and run every product in fp32.  Here
(SURVEY.md section 8f, rank 2): weights are drawn once and their int8 codes prepared once; all heads'
projections are one quantized product (qg_attention_forward); W_O and the FFN run through
qg_linear_forward_act (ReLU inside the GEMM epilogue); ADD & NORM is qg_add_layernorm_f32.  The
arithmetic of every step is the reference's (same rounding order, including its quirks: the residual
added is the attention output, transformer.cu:57,74, and the "layernorm" divides by the variance,
op_layernorm.cuh:28).  `batch` independent sequences are processed per call; the reference has no
batch dimension (batch = 1).
"""
from __future__ import annotations

import ctypes as C
import importlib

import torch

_qg = importlib.import_module(__name__.rsplit(".", 1)[0])


def add_layernorm(A: torch.Tensor, R, B: torch.Tensor) -> None:
    """op_add(A, R, T); op_layernorm(T, B) (transformer.cu:57-58).  R may be None; B may be A."""
    assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.shape == B.shape
    pa, lda = _qg._dev2d(A)
    pb, ldb = _qg._dev2d(B)
    pr, ldr = _qg._dev2d(R) if R is not None else (None, 0)
    _qg._check(_qg.lib().qg_add_layernorm_f32(pa, lda, pr, ldr, A.shape[0], A.shape[1], pb, ldb, _qg._stream()),
               "qg_add_layernorm_f32")


class PreparedLinear:
    """y = act(x @ w + b) with w's int8 codes prepared once.  bias=False: a bare product (W_O, transformer.cu:52-54)."""

    def __init__(self, in_dim: int, out_dim: int, bias: bool = True, device="cuda", range_: float = 127.0,
                 mode: int = _qg.MODE_REF_EXACT):
        self.in_dim, self.out_dim, self.range, self.mode = in_dim, out_dim, range_, mode
        self.w = torch.empty((in_dim, out_dim), dtype=torch.float32, device=device)
        self.b = torch.empty((1, out_dim), dtype=torch.float32, device=device) if bias else None
        self._wq = None
        self._ws = None

    def init_uniform(self, lo=None, hi=None, generator=None):
        mx = 1.0 / (self.in_dim ** 0.5)  # linear.cuh:33-39
        self.w.uniform_(-mx if lo is None else lo, mx if hi is None else hi, generator=generator)
        if self.b is not None:
            self.b.uniform_(-mx, mx, generator=generator)
        self._wq = None

    def forward(self, x: torch.Tensor, y: torch.Tensor, act: int = 0) -> None:
        if self._wq is None:
            self._wq = _qg.prepare_weights(self.w, self.range, self.mode)
        Wt, Cw = self._wq
        M, K, N = x.shape[0], self.in_dim, self.out_dim
        need = _qg.workspace_bytes(M, N, K)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        px, ldx = _qg._dev2d(x)
        pq, ldq = _qg._dev2d(Wt)
        py, ldy = _qg._dev2d(y)
        pb = _qg._vec(self.b.reshape(-1), N) if self.b is not None else None
        _qg._check(_qg.lib().qg_linear_forward_act(px, ldx, _qg._dt(x), pq, ldq, _qg._vec(Cw, N), pb, act, py, ldy, _qg._dt(y),
                                                   M, N, K, C.c_float(self.range), self.mode,
                                                   C.c_void_p(self._ws.data_ptr()), C.c_size_t(self._ws.numel()),
                                                   _qg._stream()), "qg_linear_forward_act")


ACT_NONE, ACT_RELU = 0, 1


def ffn_chain(ll1: "PreparedLinear", ll2: "PreparedLinear", x, hidden: torch.Tensor, y: torch.Tensor, xq=None, cx=None) -> None:
    """ll1.forward -> op_relu -> ll2.forward (transformer.cu:63-71) as one qg_ffn_forward call; the input either as floats
    (x) or as int8 codes + Cx (from add_layernorm_quant)."""
    for lin in (ll1, ll2):
        if lin._wq is None:
            lin._wq = _qg.prepare_weights(lin.w, lin.range, lin.mode)
    (W1t, Cw1), (W2t, Cw2) = ll1._wq, ll2._wq
    _qg.ffn_forward(x, W1t, Cw1, ll1.b, W2t, Cw2, ll2.b, hidden, y, ll1.range, ll1.mode, Xq=xq, Cx=cx)


class EncoderBlock:
    """One iteration of the Encoder loop, transformer.cu:24-76."""

    def __init__(self, d_model: int, n_heads: int, d_ff: int, device="cuda"):
        self.d_model, self.n_heads, self.d_ff = d_model, n_heads, d_ff
        self.attn = _qg.MultiHeadAttention(d_model, n_heads, device=device)
        self.W_O = PreparedLinear(d_model, d_model, bias=False, device=device)
        self.ll1 = PreparedLinear(d_model, d_ff, device=device)
        self.ll2 = PreparedLinear(d_ff, d_model, device=device)
        self._buf = None

    def init_uniform(self, generator=None):
        self.attn.init_uniform(generator)
        self.W_O.init_uniform(-1.0, 1.0, generator)  # transformer.cu:53
        self.ll1.init_uniform(generator=generator)
        self.ll2.init_uniform(generator=generator)

    def _buffers(self, T: int, device):
        if self._buf is None or self._buf[0].shape[0] != T:
            self._buf = (torch.empty((T, self.d_model), device=device), torch.empty((T, self.d_ff), device=device))
        return self._buf

    def forward(self, x: torch.Tensor, out: torch.Tensor, batch: int = 1, fused: bool = True) -> None:
        mh, ffn = self._buffers(x.shape[0], x.device)
        self.attn.forward(x, x, mh, batch=batch, prepared=True)   # :27-50 (codes of W_q | W_k | W_v made once)
        self.W_O.forward(mh, out)                  # :54
        if fused:
            # ADD & NORM emits ll1's int8 codes, ll1's epilogue the row maxima ll2's quantizer needs
            xq, cx = _qg.add_layernorm_quant(out, mh, out)            # :57-58 (+ ll1's row quantizer)
            ffn_chain(self.ll1, self.ll2, None, ffn, out, xq, cx)     # :63-71
        else:
            add_layernorm(out, mh, out)                # :57-58
            self.ll1.forward(out, ffn, ACT_RELU)       # :63-67
            self.ll2.forward(ffn, out)                 # :69-71
        add_layernorm(out, mh, out)                # :74-75


class DecoderBlock:
    """One iteration of the Decoder loop, transformer.cu:91-166."""

    def __init__(self, d_model: int, n_heads: int, d_ff: int, device="cuda"):
        self.d_model, self.n_heads, self.d_ff = d_model, n_heads, d_ff
        self.self_attn = _qg.MultiHeadAttention(d_model, n_heads, device=device)
        self.cross_attn = _qg.MultiHeadAttention(d_model, n_heads, device=device)
        self.W_O1 = PreparedLinear(d_model, d_model, bias=False, device=device)
        self.W_O2 = PreparedLinear(d_model, d_model, bias=False, device=device)
        self.ll1 = PreparedLinear(d_model, d_ff, device=device)
        self.ll2 = PreparedLinear(d_ff, d_model, device=device)
        self._buf = None

    def init_uniform(self, generator=None):
        self.self_attn.init_uniform(generator)
        self.cross_attn.init_uniform(generator)
        self.W_O1.init_uniform(-1.0, 1.0, generator)  # transformer.cu:118
        self.W_O2.init_uniform(-1.0, 1.0, generator)  # transformer.cu:143
        self.ll1.init_uniform(generator=generator)
        self.ll2.init_uniform(generator=generator)

    def forward(self, x: torch.Tensor, enc_output: torch.Tensor, out: torch.Tensor, batch: int = 1) -> None:
        T = x.shape[0]
        if self._buf is None or self._buf[0].shape[0] != T:
            self._buf = (torch.empty((T, self.d_model), device=x.device), torch.empty((T, self.d_ff), device=x.device))
        mh, ffn = self._buf
        self.self_attn.forward(x, x, mh, batch=batch, prepared=True)              # :99-116
        self.W_O1.forward(mh, out)                                 # :119
        add_layernorm(out, mh, out)                                # :123-124
        self.cross_attn.forward(out, enc_output, mh, batch=batch, prepared=True)  # :127-140 (queries from the decoder, keys/values from the encoder)
        self.W_O2.forward(mh, out)                                 # :144
        xq, cx = _qg.add_layernorm_quant(out, mh, out)             # :148-149 (+ ll1's row quantizer)
        ffn_chain(self.ll1, self.ll2, None, ffn, out, xq, cx)      # :154-162
        add_layernorm(out, mh, out)                                # :165-166


class Encoder:
    def __init__(self, d_model: int, n_heads: int, n_blocks: int, d_ff: int, device="cuda"):
        self.blocks = [EncoderBlock(d_model, n_heads, d_ff, device) for _ in range(n_blocks)]

    def init_uniform(self, generator=None):
        for b in self.blocks:
            b.init_uniform(generator)

    def forward(self, x: torch.Tensor, out: torch.Tensor, batch: int = 1) -> None:
        cur = x
        for blk in self.blocks:  # block 0 reads X, later blocks read the running output (transformer.cu:35-39);
            blk.forward(cur, out, batch)  # in place is safe: the attention has consumed its input before `out` is written
            cur = out


class Decoder:
    def __init__(self, d_model: int, n_heads: int, n_blocks: int, d_ff: int, device="cuda"):
        self.blocks = [DecoderBlock(d_model, n_heads, d_ff, device) for _ in range(n_blocks)]

    def init_uniform(self, generator=None):
        for b in self.blocks:
            b.init_uniform(generator)

    def forward(self, x: torch.Tensor, enc_output: torch.Tensor, out: torch.Tensor, batch: int = 1) -> None:
        cur = x
        for blk in self.blocks:
            blk.forward(cur, enc_output, out, batch)
            cur = out
