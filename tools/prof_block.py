"""One encoder block and one decoder block of config 3 (32 x 128 tokens, d_model 512, 8 heads, d_ff 2048), run three times each:
the launch sequence for `ncu --metrics gpu__time_duration.sum` (read the last repetition)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
tf = importlib.import_module(qg.__name__ + ".transformer")
DEV = "cuda"
batch, seq, d_model, heads, d_ff = 32, 128, 512, 8, 2048
T = batch * seq
g = torch.Generator(device=DEV).manual_seed(0)
enc = tf.EncoderBlock(d_model, heads, d_ff, DEV); enc.init_uniform(g)
dec = tf.DecoderBlock(d_model, heads, d_ff, DEV); dec.init_uniform(g)
x = torch.randn((T, d_model), device=DEV, generator=g)
e = torch.randn((T, d_model), device=DEV, generator=g)
out = torch.empty((T, d_model), device=DEV)
for rep in range(3):
    enc.forward(x, out, batch)
torch.cuda.synchronize()
for rep in range(3):
    dec.forward(x, e, out, batch)
torch.cuda.synchronize()
print("ok")
