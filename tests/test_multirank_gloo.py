"""N>1 path on CPU: two gloo ranks shard a batch by cells, each computes its shard (with the
oracle standing in for the device kernels — this test is about the sharding / assembly logic),
all-gather the row lengths and rebuild the global CSR index; must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_cells, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from alevin_fry_b200 import QuantOpts, shard
    import synth
    spec = synth.SynthSpec(reads_mean=150.0, n_genes=500)
    t2g = synth.tid_to_gid(spec)
    opts = QuantOpts(resolution="cr-like", num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows)
    first, cnt = shard.shard_range(n_cells, rank, world)
    mine = oracle_lib.oracle_quant(opts, t2g, synth.generate(spec, first, cnt, n_threads=2), n_threads=2)
    max_cells = max(shard.shard_range(n_cells, r, world)[1] for r in range(world))
    lens = shard.gather_row_lengths(torch.from_numpy(mine.num_expr.astype(np.int32)), max_cells)
    rp = shard.global_row_ptr(lens, n_cells, world)
    # every rank's own rows land at rp[first : first+cnt+1]
    assert torch.equal(rp[first:first + cnt + 1] - rp[first], torch.from_numpy(mine.row_ptr.astype(np.int64)))
    # ... and the payload: one matrix on every rank (variable-length all-gather)
    rp2, cols, vals = shard.assemble_csr(torch.from_numpy(mine.num_expr.astype(np.int32)), torch.from_numpy(mine.col.astype(np.int32)),
                                         torch.from_numpy(mine.val), n_cells)
    whole = oracle_lib.oracle_quant(opts, t2g, synth.generate(spec, 0, n_cells, n_threads=2), n_threads=2)
    assert torch.equal(rp2, rp)
    assert np.array_equal(cols.numpy().astype(np.uint32), whole.col) and np.array_equal(vals.numpy(), whole.val)
    # the form bench.py times (compact=False): the gathered payload stays padded to the largest rank, rank r's entries at
    # r * stride; staging arrays longer than the stride are sent as views (no padded copy), as the device API's are
    cap = int(mine.nnz) + 1000 + 17 * rank
    col_cap = torch.full((cap,), -7, dtype=torch.int32); col_cap[:mine.nnz] = torch.from_numpy(mine.col.astype(np.int32))
    val_cap = torch.full((cap,), -7.0, dtype=torch.float32); val_cap[:mine.nnz] = torch.from_numpy(mine.val)
    rp3, ac, av, stride, nnz_r = shard.assemble_csr(torch.from_numpy(mine.num_expr.astype(np.int32)), col_cap, val_cap, n_cells, compact=False)
    assert torch.equal(rp3, rp) and int(nnz_r.sum()) == int(whole.nnz) and ac.numel() == world * stride
    got_c = torch.cat([ac[r * stride: r * stride + int(nnz_r[r])] for r in range(world)])
    got_v = torch.cat([av[r * stride: r * stride + int(nnz_r[r])] for r in range(world)])
    assert np.array_equal(got_c.numpy().astype(np.uint32), whole.col) and np.array_equal(got_v.numpy(), whole.val)
    if rank == 0:
        out.put((rp.numpy().tolist() == whole.row_ptr.astype(np.int64).tolist(), int(rp[-1]), int(whole.nnz)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_cell_sharding_assembles_the_global_index():
    from alevin_fry_b200 import shard
    assert [shard.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 3), (7, 3)]
    assert sum(shard.shard_range(1_000_003, r, 8)[1] for r in range(8)) == 1_000_003
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 41, out)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    ok, nnz_a, nnz_b = out.get(timeout=5)
    assert ok and nnz_a == nnz_b
