"""Megatron pairing of the FFN over the GPUs of one box.

The reference's FFN is `ll1.forward -> op_relu -> ll2.forward` (src/transformer.cu:63-71) and has no multi-GPU
code.  Column-parallel alone (colpar.py) must all-gather fc1's [T, d_ff] output before fc2 can run; here fc1 is
column-parallel and fc2 row-parallel on the same slice of the hidden features, so the activation never leaves
the GPU that produced it and the only exchange is a reduce-scatter (plus, optionally, an all-gather) of fc2's
[T, d_out] partial products -- 4x less data for d_ff = 4 d_model, moved by the kernels themselves:

  * the fc2 GEMM's epilogue stores column block b of the rank's partial product straight into slot `rank` of the
    GPU that owns block b (TMA stores over NVLink into symmetric memory: qg_ffn_forward_rowpar);
  * after a barrier, qg_reduce_partials adds the P slots in ascending rank order (+ bias) and writes the owner's
    block to its own result and to the same block of every peer's (plain 16-byte stores over NVLink).

Numerics: fc2's row / column scales are those of the rank's slice, so the result differs in the last bits from
the single-GPU layer -- an explicitly specified mode of its own (include/qgemm.h), bit-identical to
oracle.megatron_ffn for every P, and independent of timing (fixed summation order).

`exchange="collective"` swaps the kernel-carried exchange for torch.distributed collectives (all_to_all of the
column blocks, ordered local sum, all_gather): the baseline the fused path is measured against, and -- with
`compute` injected -- the form the host logic is tested in under gloo on CPU.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from .colpar import shard_bounds


def block_cols_for(d_out: int, world: int, align: int = 64) -> int:
    """Width of one owner's column block of the reduce-scatter: equal blocks, multiples of `align` columns (one TMA
    store of the scattering epilogue carries 32 fp32 / 64 16-bit columns and must not straddle two owners)."""
    per = (d_out + world - 1) // world
    return (per + align - 1) // align * align


class MegatronFFN:
    """y = relu(x @ W1 + b1) @ W2 + b2 with the d_ff hidden features sliced over `world` ranks.

    forward(x) returns y [M, d_out] on every rank (gather=True) or this rank's column block [M, hi-lo] (gather=False).
    compute(x, w1_p, b1_p, w2_p) -> this rank's partial product [M, d_out] replaces the CUDA path (CPU tests)."""

    def __init__(self, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor, b2: Optional[torch.Tensor], rank: int,
                 world: int, group=None, exchange: str = "fused", h_dtype: torch.dtype = torch.float32,
                 part_dtype: torch.dtype = torch.float32, out_dtype: torch.dtype = torch.float32, gather: bool = True,
                 compute: Optional[Callable] = None, range_: float = 127.0, mode: int = 0, align: int = 16, chunks: int = 0,
                 comm_sms: int = 0, gather_engine: str = "kernel"):
        assert exchange in ("fused", "collective")
        # fused exchange: the tokens are cut into `chunks` row blocks; block c's reduce + gather (side stream) runs under block
        # c+1's GEMMs.  0 = choose (4 blocks of >= 512 rows when the batch allows it).  Rows are independent, so the result does
        # not depend on the chunking.
        self.chunks = chunks
        # SMs kept free of the (persistent) GEMMs while the tokens are processed in several row blocks, so that the side
        # stream's barrier / reduce / gather kernels of block c actually run under block c+1's products instead of queueing
        # behind them -- the same reason communication libraries reserve SMs for their channels
        # (measured: the reservation costs more than it hides -- 140 instead of 148 SMs turns 72-tile products from one wave into
        # two; P = 2: 991 -> 1220 us with 8 SMs held back.  Kept as a knob, default 0.)
        self.comm_sms = comm_sms
        # the gather of the reduced blocks: "kernel" = peer stores from the reduce kernel (default); "copy" = cudaMemcpy2DAsync on
        # the copy engines, which can run under the next row block's GEMMs although those hold every SM -- measured slower all
        # the same (P = 8, bf16 partials: 672 vs 505 us with 2 row blocks, 723 vs 537 us with 1: seven strided 2-D copies of
        # 2304-byte rows per rank are a poor load for the engines); "auto" = copy when row blocks pipeline
        assert gather_engine in ("auto", "kernel", "copy")
        self.gather_engine = gather_engine
        self.rank, self.world, self.group = rank, world, group
        self.exchange, self.gather = exchange, gather
        self.h_dtype, self.part_dtype, self.out_dtype = h_dtype, part_dtype, out_dtype
        self.range, self.mode = range_, mode
        self.d_in, self.d_ff = w1.shape
        self.d_out = w2.shape[1]
        assert w2.shape[0] == self.d_ff
        self.flo, self.fhi = shard_bounds(self.d_ff, world, rank, align)   # this rank's hidden features
        self.bc = block_cols_for(self.d_out, world)                       # owner b holds columns [b*bc, min((b+1)*bc, d_out))
        self.olo, self.ohi = min(rank * self.bc, self.d_out), min((rank + 1) * self.bc, self.d_out)
        self.w1 = w1[:, self.flo:self.fhi].contiguous()
        self.b1 = None if b1 is None else b1.reshape(-1)[self.flo:self.fhi].contiguous().float()
        self.w2 = w2[self.flo:self.fhi, :].contiguous()
        self.b2 = None if b2 is None else b2.reshape(-1).contiguous().float()
        self._compute = compute
        self._prepared = False
        self._m = None

    # ---- CUDA state -------------------------------------------------------------------------
    def _prepare(self, device):
        from . import prepare_weights

        self.w1t, self.cw1 = prepare_weights(self.w1.to(device), self.range, self.mode)
        self.w2t, self.cw2 = prepare_weights(self.w2.to(device), self.range, self.mode)
        self.b1 = None if self.b1 is None else self.b1.to(device)
        self.b2 = None if self.b2 is None else self.b2.to(device)
        self._prepared = True

    def _ensure(self, m: int, device):
        if self._m == m:
            return
        from . import ffn_workspace_bytes

        self.h = torch.empty((m, self.fhi - self.flo), dtype=self.h_dtype, device=device)
        self.ws = torch.empty(max(ffn_workspace_bytes(m, self.d_in, self.fhi - self.flo, self.d_out), 256) + 256,
                              dtype=torch.uint8, device=device)
        off = (-self.ws.data_ptr()) % 256
        self.ws = self.ws[off:]
        if self.exchange == "fused" and self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem

            grp = self.group if self.group is not None else dist.group.WORLD
            self.slots = symm_mem.empty((self.world, m, self.bc), dtype=self.part_dtype, device=device)
            self.slots_hdl = symm_mem.rendezvous(self.slots, grp)
            es = self.slots.element_size()
            # owner b's slot for THIS rank's partial product
            self.part_ptrs = [int(self.slots_hdl.buffer_ptrs[b]) + self.rank * m * self.bc * es for b in range(self.world)]
            if self.gather:
                self.out = symm_mem.empty((m, self.d_out), dtype=self.out_dtype, device=device)
                self.out_hdl = symm_mem.rendezvous(self.out, grp)
                eo = self.out.element_size()
                self.peer_out = [int(self.out_hdl.buffer_ptrs[r]) + self.olo * eo for r in range(self.world) if r != self.rank]
                import os

                mc = int(getattr(self.out_hdl, "multicast_ptr", 0) or 0)  # NVSwitch multicast mapping of the result, if any
                self.out_mc = mc + self.olo * eo if mc and os.environ.get("QG_MULTICAST") == "1" else 0  # opt-in: see colpar.py
            else:
                self.out = torch.empty((m, self.ohi - self.olo), dtype=self.out_dtype, device=device)
                self.peer_out = []
                self.out_mc = 0
            self._fresh = True
        else:
            self.slots = torch.empty((max(self.world, 1), m, self.bc), dtype=self.part_dtype, device=device)
            self.part_ptrs = [self.slots.data_ptr() + b * m * self.bc * self.slots.element_size() for b in range(self.world)]
        self._m = m

    # ---- this rank's partial product, as [world, M, bc] column blocks -------------------------
    def _partial_blocks(self, x: torch.Tensor) -> torch.Tensor:
        """Collective / CPU form: the rank's whole partial product cut into the owners' blocks (zero padded)."""
        m = x.shape[0]
        if self._compute is not None:
            part = self._compute(x, self.w1, self.b1, self.w2)
            blocks = torch.zeros((self.world, m, self.bc), dtype=part.dtype, device=part.device)
            for b in range(self.world):
                lo, hi = min(b * self.bc, self.d_out), min((b + 1) * self.bc, self.d_out)
                blocks[b, :, : hi - lo] = part[:, lo:hi]
            return blocks
        from . import ffn_forward_rowpar

        if not self._prepared:
            self._prepare(x.device)
        self._ensure(m, x.device)
        # destination b = block b of a local [world, M, bc] buffer: the same scattering epilogue, no peers involved
        ffn_forward_rowpar(x, self.w1t, self.cw1, self.b1, self.w2t, self.cw2, self.h, self.part_ptrs, self.bc, self.bc,
                           self.part_dtype, self.d_out, self.range, self.mode, workspace=self.ws)
        return self.slots

    def _reduce_ordered(self, slots: torch.Tensor) -> torch.Tensor:
        """((s_0 + s_1) + ...) + b2[block] in fp32, ascending rank order, rounded to the output dtype (host-side form)."""
        n = self.ohi - self.olo
        acc = slots[0, :, :n].float()
        for p in range(1, slots.shape[0]):
            acc = acc + slots[p, :, :n].float()
        if self.b2 is not None:
            acc = acc + self.b2[self.olo:self.ohi].to(acc.device).reshape(1, -1)
        return acc.to(self.out_dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.exchange == "fused" and self.world > 1 and self._compute is None:
            return self._forward_fused(x)
        blocks = self._partial_blocks(x)
        if self.world == 1:
            return self._reduce_ordered(blocks[:1])
        recv = torch.empty_like(blocks)  # recv[p] = rank p's partial of MY block
        if blocks.is_cuda:
            dist.all_to_all_single(recv, blocks.contiguous(), group=self.group)
        else:  # gloo: no all_to_all; gather everything and keep my block (CPU tests only)
            everything = [torch.empty_like(blocks) for _ in range(self.world)]
            dist.all_gather(everything, blocks.contiguous(), group=self.group)
            recv = torch.stack([everything[p][self.rank] for p in range(self.world)])
        mine = self._reduce_ordered(recv)
        if not self.gather:
            return mine
        m = x.shape[0]
        pad = torch.zeros((m, self.bc), dtype=mine.dtype, device=mine.device)
        pad[:, : mine.shape[1]] = mine
        allb = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(allb, pad, group=self.group)
        out = torch.empty((m, self.d_out), dtype=mine.dtype, device=mine.device)
        for b in range(self.world):
            lo, hi = min(b * self.bc, self.d_out), min((b + 1) * self.bc, self.d_out)
            out[:, lo:hi] = allb[b][:, : hi - lo]
        return out

    def _row_blocks(self, m: int):
        c = self.chunks
        if c <= 0:
            # measured on the OPT-66B FFN, T = 4096 (profiles/r2_megatron_opt66b_p{2,8}_row_blocks_run2{0,1}.json): at P = 2 the
            # exchange is a small share and cutting the GEMMs costs more than it hides (1004 -> 1079 us with 4 blocks); at P = 8
            # fp32 partials gain from 4 blocks (772 -> 659 us), 16-bit partials from 2 (539 -> 505 us), 8 blocks always lose
            if self.world <= 2 or m < 2048:
                c = 1
            else:
                c = 4 if self.part_dtype == torch.float32 else 2
        c = max(1, min(c, m // 256 if m >= 256 else 1))
        step = -(-m // c)
        step = -(-step // 256) * 256  # whole 2-SM tiles per block
        return [(r0, min(r0 + step, m)) for r0 in range(0, m, step)]

    def _forward_fused(self, x: torch.Tensor) -> torch.Tensor:
        from . import copy_2d_async, ffn_forward_rowpar, reduce_partials

        if not self._prepared:
            self._prepare(x.device)
        m = x.shape[0]
        self._ensure(m, x.device)
        blocks = self._row_blocks(m)
        main = torch.cuda.current_stream()
        from . import device_info, set_gemm_sm_limit

        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=x.device, priority=-1)
        side = self._side
        reserve = self.comm_sms if len(blocks) > 1 else 0
        if reserve:
            set_gemm_sm_limit(device_info()[0] - reserve)
        # (a) every peer has finished reducing the previous contents of its slots (and reading its previous result)
        if self._fresh or not self.gather:
            self.slots_hdl.barrier(channel=0)
            self._fresh = False
        n = self.ohi - self.olo
        es, eo = self.slots.element_size(), self.out.element_size()
        bias = None if self.b2 is None else self.b2[self.olo:self.ohi]
        side.wait_stream(main)
        for r0, r1 in blocks:
            ptrs = [p + r0 * self.bc * es for p in self.part_ptrs]
            ffn_forward_rowpar(x[r0:r1], self.w1t, self.cw1, self.b1, self.w2t, self.cw2, self.h[r0:r1], ptrs, self.bc, self.bc,
                               self.part_dtype, self.d_out, self.range, self.mode, workspace=self.ws)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(side):  # this block's exchange tail runs under the next block's GEMMs
                side.wait_event(done)
                # (b) every rank's partial products of this row block have landed in their owners' slots
                self.slots_hdl.barrier(channel=1)
                if n > 0:
                    own = self.out[r0:r1, self.olo:self.ohi] if self.gather else self.out[r0:r1]
                    ldo = self.out.stride(0)
                    peers = [p + r0 * ldo * eo for p in self.peer_out]
                    by_copy = self.gather and bool(peers) and (self.gather_engine == "copy" or
                                                               (self.gather_engine == "auto" and len(blocks) > 1))
                    reduce_partials(self.slots[:, r0:r1, :], bias, own, [] if by_copy else peers, n,
                                    0 if by_copy else (self.out_mc + r0 * ldo * eo if self.out_mc else 0),
                                    max_ctas=8 * reserve if reserve else 0)
                    if by_copy:  # the owner's block to the same place in every peer's result, by the copy engines
                        for pdst in peers:
                            copy_2d_async(pdst, ldo * eo, own.data_ptr(), ldo * eo, n * eo, r1 - r0)
        with torch.cuda.stream(side):
            if self.gather:
                # (c) every block of every rank's result is in place; it also orders the next forward's stores after this reduce
                self.out_hdl.barrier(channel=0)
        main.wait_stream(side)
        if reserve:
            set_gemm_sm_limit(0)
        return self.out
