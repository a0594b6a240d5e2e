"""Per-op timing of one config-3 encoder block (32 x 128 tokens, d_model 512, 8 heads, d_ff 2048)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
tf = importlib.import_module(qg.__name__ + ".transformer")
from tools.bench_configs import timed
DEV = "cuda"
batch, seq, d_model, heads, d_ff = 32, 128, 512, 8, 2048
T = batch * seq
g = torch.Generator(device=DEV).manual_seed(0)
blk = tf.EncoderBlock(d_model, heads, d_ff, DEV); blk.init_uniform(g)
x = torch.randn((T, d_model), device=DEV, generator=g)
mh = torch.empty((T, d_model), device=DEV); out = torch.empty((T, d_model), device=DEV); ffn = torch.empty((T, d_ff), device=DEV)
res = {}
res["mha"] = timed(lambda i: blk.attn.forward(x, x, mh, batch=batch))
res["W_O"] = timed(lambda i: blk.W_O.forward(mh, out))
res["addnorm"] = timed(lambda i: tf.add_layernorm(out, mh, out))
res["ll1_relu"] = timed(lambda i: blk.ll1.forward(out, ffn, tf.ACT_RELU))
res["ll2"] = timed(lambda i: blk.ll2.forward(ffn, out))
res["block"] = timed(lambda i: blk.forward(x, out, batch))
# attention pieces
nq = heads * 64
proj = torch.empty((T, 3 * nq), device=DEV)
res["attn_projection"] = timed(lambda i: qg.op_quantized_mm(x, blk.attn.W_qkv, proj, 127.0))
S = torch.empty((batch * heads * seq, seq), device=DEV)
res["softmax_32768x128"] = timed(lambda i: qg.op_softmax(S, S, 0.125))
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
