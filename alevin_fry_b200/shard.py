"""Multi-GPU plumbing for the cell-sharded quant path.

Cells are independent work items (reference src/quant.rs:1389, 735), so ranks take contiguous
cell ranges and never exchange data during compute. The collectives belong to the ASSEMBLY of the one
output matrix (reference: every worker's triplets end up in one TriMat, src/quant.rs:1811-1847): an
all-gather of per-cell row lengths, from which every rank derives the global CSR row pointer, and a
variable-length all-gather of the (col, val) payload. Works with NCCL (CUDA tensors, NVLink / NVSwitch)
and gloo (CPU tensors)."""
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


def assemble_csr(num_expr: torch.Tensor, col: torch.Tensor, val: torch.Tensor, n_cells: int, group=None, scratch=None, compact=True):
    """The one sparse matrix of the job on every rank: (row_ptr int64 [n_cells+1], col int32 [nnz], val float32 [nnz]).

    Rank r holds the rows of its cell range (shard_range) as CSR pieces `num_expr` (row lengths) and `col` / `val`
    (only the first sum(num_expr) entries are used). Two collectives: row lengths (padded to the largest range), then
    the payload padded to the largest per-rank nnz — one all_gather_into_tensor each for col and val, so that NCCL
    moves two large messages per rank over NVLink instead of many small ones. `scratch` (a dict) keeps the padded
    buffers between calls. compact=False skips the final concatenation and returns the gathered, padded payload
    (rank r's entries start at r * stride): (row_ptr, col_padded, val_padded, stride, nnz_per_rank)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = num_expr.device
    _, mine = shard_range(n_cells, rank, world)
    max_cells = max(shard_range(n_cells, r, world)[1] for r in range(world))
    lens = gather_row_lengths(num_expr[:mine], max_cells, group)
    nnz_r = lens.to(torch.int64).sum(dim=1)                      # [world] (padding rows are zero)
    max_nnz = int(nnz_r.max().item())
    my_nnz = int(nnz_r[rank].item())
    scratch = scratch if scratch is not None else {}
    key = (max_nnz, world, str(dev))
    m = max(max_nnz, 1)
    if scratch.get("key") != key:
        scratch["key"] = key
        scratch["pc"] = scratch["pv"] = None
        scratch["ac"] = torch.empty(world * m, dtype=torch.int32, device=dev)
        scratch["av"] = torch.empty(world * m, dtype=torch.float32, device=dev)
    ac, av = scratch["ac"], scratch["av"]
    # every rank sends `m` entries. The staging arrays of the device API hold one slot per alignment (>= any rank's nnz in
    # practice), so the send buffer is normally a VIEW of col / val — what lies behind a rank's own nnz is never read by
    # anyone; only arrays shorter than m are copied into a padded buffer first.
    if col.dtype == torch.int32 and col.is_contiguous() and col.numel() >= m:
        pc = col[:m]
    else:
        if scratch["pc"] is None:
            scratch["pc"] = torch.zeros(m, dtype=torch.int32, device=dev)
        pc = scratch["pc"]
        pc[:my_nnz] = col[:my_nnz].to(torch.int32)
    if val.dtype == torch.float32 and val.is_contiguous() and val.numel() >= m:
        pv = val[:m]
    else:
        if scratch["pv"] is None:
            scratch["pv"] = torch.zeros(m, dtype=torch.float32, device=dev)
        pv = scratch["pv"]
        pv[:my_nnz] = val[:my_nnz]
    dist.all_gather_into_tensor(ac, pc, group=group)
    dist.all_gather_into_tensor(av, pv, group=group)
    rp = global_row_ptr(lens, n_cells, world)
    if not compact:
        return rp, ac, av, m, nnz_r
    cols = torch.cat([ac[r * m: r * m + int(nnz_r[r].item())] for r in range(world)])
    vals = torch.cat([av[r * m: r * m + int(nnz_r[r].item())] for r in range(world)])
    return rp, cols, vals
