"""Multi-GPU plumbing for the cell-sharded quant path.

Cells are independent work items (reference src/quant.rs:1389, 735), so ranks take contiguous
cell ranges and never exchange data during compute. The only collective is an all-gather of
per-cell row lengths, from which every rank derives the global CSR row pointer / its own row
offset in the global matrix. Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_cells: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range of cells for `rank`: (first, count)."""
    base, rem = divmod(n_cells, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_row_lengths(num_expr: torch.Tensor, max_cells: int, group=None) -> torch.Tensor:
    """All-gather per-cell row lengths. Ranks may hold different cell counts: each pads to
    `max_cells` (>= every rank's count) with zeros. Returns [world, max_cells] int32."""
    world = dist.get_world_size(group)
    padded = torch.zeros(max_cells, dtype=torch.int32, device=num_expr.device)
    padded[: num_expr.numel()] = num_expr.to(torch.int32)
    out = torch.empty(world * max_cells, dtype=torch.int32, device=num_expr.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out.view(world, max_cells)


def global_row_ptr(lengths: torch.Tensor, n_cells: int, world: int) -> torch.Tensor:
    """Global CSR row pointer (int64, [n_cells+1]) from gathered [world, max_cells] lengths."""
    parts = []
    for r in range(world):
        _, cnt = shard_range(n_cells, r, world)
        parts.append(lengths[r, :cnt].to(torch.int64))
    allc = torch.cat(parts)
    rp = torch.zeros(n_cells + 1, dtype=torch.int64, device=lengths.device)
    torch.cumsum(allc, 0, out=rp[1:])
    return rp
