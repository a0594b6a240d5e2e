"""BASELINE config 3 (6 + 6 block stack, 32 x 128 tokens): the whole forward as ONE CUDA graph against call-by-call launches.

Every kernel of the stack goes through the C ABI on torch's current stream with caller-owned buffers, no allocation and no
synchronisation inside a forward, so the stream can be captured (programmatic-dependent-launch edges included).  The block
kernels are 5-45 us each: replaying the captured graph removes the host-side cost of ~130 C-ABI calls per forward.
Prints eager and graph times and checks that the graph's output equals the eager one bit for bit.
"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
tf = importlib.import_module(qg.__name__ + ".transformer")
DEV = torch.device("cuda", 0)
batch, seq, d_model, heads, d_ff, n_blocks = 32, 128, 512, 8, 2048, 6
T = batch * seq
g = torch.Generator(device=DEV).manual_seed(0)
enc, dec = tf.Encoder(d_model, heads, n_blocks, d_ff, DEV), tf.Decoder(d_model, heads, n_blocks, d_ff, DEV)
enc.init_uniform(g); dec.init_uniform(g)
X = torch.randn((T, d_model), device=DEV, generator=g)
Y = torch.randn((T, d_model), device=DEV, generator=g)
eo, do = torch.empty((T, d_model), device=DEV), torch.empty((T, d_model), device=DEV)


def step():
    enc.forward(X, eo, batch)
    dec.forward(Y, eo, do, batch)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):  # warm up on the stream that will be captured (per-stream scratch is created on first use)
    for _ in range(3):
        step()
    side.synchronize()
    eager_us = timed(step)
    ref = do.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step()
    do.zero_()
    graph.replay()
    side.synchronize()
    same = bool(torch.equal(do.view(torch.int32), ref.view(torch.int32)))
    graph_us = timed(graph.replay)
print(json.dumps({"name": "transformer_enc6_dec6_b32_s128_d512_h8_ff2048", "tokens": T, "eager_us": eager_us, "graph_us": graph_us,
                  "eager_tok_per_s": T / eager_us * 1e6, "graph_tok_per_s": T / graph_us * 1e6, "graph_equals_eager_bits": same}))
assert same
