"""Column-parallel quantized linear over the GPUs of one box.

The reference has no multi-GPU code; this is the sharding BASELINE.json's multi-GPU configs name.
Column scale Cw[j] and every output column O[:, j] depend only on W[:, j] and the whole X, so W
splits along N with no cross-GPU arithmetic: rank p owns W[:, lo:hi], quantizes the replicated X
locally, computes O[:, lo:hi], and the only exchange is the gather of the output blocks.  Results
are bit-identical to the single-GPU op.

Two exchange paths:
  * "nccl"  : all_gather_into_tensor of the [M, N/P] blocks, then one permute into [M, N];
  * "fused" : the GEMM epilogue TMA-stores every output tile straight into the [M, N] buffer of
              every peer (symmetric memory over NVLink), so the transfer overlaps the main loop
              and no permute pass exists (qg_gemm_s8_dequant_multi).

`compute` is injectable so that the host-side logic (shard bounds, gather layout) is testable with
the gloo backend on CPU, where tests plug in the CPU oracle.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    """Contiguous column range [lo, hi) of `rank`; shard sizes differ by at most `align` columns
    and every boundary is a multiple of `align` (TMA wants 16-byte aligned column offsets)."""
    assert 0 <= rank < world and n >= 0 and align >= 1
    units = (n + align - 1) // align
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(lo_u * align, n), min(hi_u * align, n)


def gather_columns(local: torch.Tensor, n: int, world: int, rank: int, group=None, align: int = 1,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather column blocks [M, hi-lo] into a row-major [M, n] tensor on every rank.
    Blocks may have different widths (n not divisible by world): they are padded to the widest."""
    m = local.shape[0]
    widths = [hi - lo for lo, hi in (shard_bounds(n, world, r, align) for r in range(world))]
    wmax = max(widths)
    send = local
    if local.shape[1] != wmax:
        send = torch.zeros((m, wmax), dtype=local.dtype, device=local.device)
        send[:, : local.shape[1]] = local
    buf = torch.empty((world, m, wmax), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(buf, send.contiguous(), group=group)
    else:  # gloo (CPU tests) has no flat all-gather
        dist.all_gather([buf[r] for r in range(world)], send.contiguous(), group=group)
    if out is None:
        out = torch.empty((m, n), dtype=local.dtype, device=local.device)
    for r in range(world):
        lo, hi = shard_bounds(n, world, r, align)
        out[:, lo:hi] = buf[r, :, : hi - lo]
    return out


class ColumnParallelLinear:
    """y = x @ W + b with W [K, N] sharded by columns over `world` ranks.

    compute(x, w_shard, bias_shard) -> [M, hi-lo] is the per-rank quantized linear; the default
    runs the C-ABI path (quantize once, cached int8 shard) on the current CUDA device.
    """

    def __init__(self, w_full: torch.Tensor, bias: Optional[torch.Tensor], rank: int, world: int, group=None,
                 compute: Optional[Callable] = None, align: int = 16, range_: float = 127.0, mode: int = 0):
        self.rank, self.world, self.group, self.align = rank, world, group, align
        self.k, self.n = w_full.shape
        self.lo, self.hi = shard_bounds(self.n, world, rank, align)
        self.w = w_full[:, self.lo:self.hi].contiguous()
        self.b = None if bias is None else bias.reshape(-1)[self.lo:self.hi].contiguous()
        self.range, self.mode = range_, mode
        self._compute = compute
        self._cache = None

    def local_forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._compute is not None:
            return self._compute(x, self.w, self.b)
        from . import LinearLayer  # product path: C ABI on the current device

        if self._cache is None:
            lin = LinearLayer(self.k, self.hi - self.lo, device=x.device, dtype=self.w.dtype, range_=self.range,
                              mode=self.mode)
            lin.w = self.w.to(x.device)
            lin.b = (torch.zeros(self.hi - self.lo, device=x.device) if self.b is None else self.b.to(x.device)).reshape(1, -1).float()
            lin.quantize_weights()
            self._cache = lin
        y = torch.empty((x.shape[0], self.hi - self.lo), dtype=x.dtype, device=x.device)
        self._cache.forward(x, y)
        return y

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = self.local_forward(x)
        if self.world == 1:
            return y
        return gather_columns(y, self.n, self.world, self.rank, self.group, self.align, out)


class FusedColumnParallelLinear:
    """Column-parallel linear whose output gather is fused into the GEMM epilogue: every rank's
    [M, N] result lives in symmetric memory (torch.distributed._symmetric_memory: peer-mapped over
    NVLink), and each rank's epilogue TMA-stores its column block into all of them while its main
    loop is still running.  Two cross-GPU barriers bracket a forward (peers are done reading the
    previous result / every block has landed).  On an NVSwitch fabric the stores go to the allocation's MULTICAST address
    instead (multimem.st): the switch replicates them, the sender's egress is 1x its block instead of (P-1)x."""

    def __init__(self, w_full: torch.Tensor, bias: Optional[torch.Tensor], rank: int, world: int, group=None,
                 align: int = 16, range_: float = 127.0, mode: int = 0, multicast: Optional[bool] = None):
        from . import prepare_weights
        import os

        # multicast=None: unicast TMA stores unless QG_MULTICAST=1.  The multicast store cuts the sender's egress to 1x, but an
        # all-gather is bound by what every GPU must RECEIVE, and measured it is slower (8 GPUs, 4096^3 per GPU: 814.6 vs 775.5 us)
        self.multicast = (os.environ.get("QG_MULTICAST") == "1") if multicast is None else bool(multicast)

        self.rank, self.world, self.align = rank, world, align
        self.group = group if group is not None else dist.group.WORLD
        self.k, self.n = w_full.shape
        self.lo, self.hi = shard_bounds(self.n, world, rank, align)
        self.w = w_full[:, self.lo:self.hi].contiguous()
        self.b = None if bias is None else bias.reshape(-1)[self.lo:self.hi].contiguous().float()
        self.range, self.mode = range_, mode
        self.wt, self.cw = prepare_weights(self.w, range_, mode)
        self.out = None
        self.hdl = None
        self._xq = self._cx = None

    def _ensure_out(self, m: int, dtype, device):
        if self.out is not None and self.out.shape[0] == m and self.out.dtype == dtype:
            return
        import torch.distributed._symmetric_memory as symm_mem

        self.out = symm_mem.empty((m, self.n), dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.out, self.group)
        es = self.out.element_size()
        self.peer_ptrs = [int(self.hdl.buffer_ptrs[r]) + self.lo * es for r in range(self.world) if r != self.rank]
        # NVSwitch multicast mapping of the same allocation (0 when the fabric has none): one store reaches every GPU
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        ok = mc != 0 and self.multicast and ((self.hi - self.lo) * es) % 16 == 0 and (self.lo * es) % 16 == 0 and (self.n * es) % 16 == 0
        self.mc_ptr = mc + self.lo * es if ok else 0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import absmax_quant_rows, gemm_s8_dequant_ex, gemm_s8_dequant_mc

        m = x.shape[0]
        self._ensure_out(m, x.dtype, x.device)
        if self._xq is None or self._xq.shape[0] != m:
            ldk = (self.k + 15) // 16 * 16  # TMA wants 16-byte aligned rows
            self._xq = torch.zeros((m, ldk), dtype=torch.int8, device=x.device)[:, : self.k]
            self._cx = torch.empty(m, dtype=torch.float32, device=x.device)
        absmax_quant_rows(x, self.range, self.mode, self._xq, self._cx)
        self.hdl.barrier(channel=0)  # every peer has finished with the previous contents of its matrix
        if self.mc_ptr:
            gemm_s8_dequant_mc(self._xq, self.wt, True, self._cx, self.cw, self.out[:, self.lo:self.hi], self.mc_ptr,
                               self.range, self.b)
        else:
            gemm_s8_dequant_ex(self._xq, self.wt, True, self._cx, self.cw, self.out[:, self.lo:self.hi], self.peer_ptrs,
                               self.range, self.b)
        self.hdl.barrier(channel=1)  # all blocks of all ranks have landed
        return self.out
