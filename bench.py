#!/usr/bin/env python
"""bench.py — headline benchmark of the quant hot path (BASELINE.json metric).

A "step" is one pass of the hot path (per-cell UMI resolution -> sparse cell x gene rows)
over the whole per-GPU synthetic workload. Workload at N=1 = BASELINE.json configs[1] (C2):
100k cells x ~200M reads, 10x-v3 (16 bp BC / 12 bp UMI), `--resolution cr-like`. Under
torchrun every rank generates and processes its own C2-sized shard of cells (weak scaling,
cells are independent: no data-path collective; the only exchange is the per-rank row-count
all-gather that assembles the global matrix index).

  value : cells/s, inputs resident in HBM when the timed region starts (CUDA events on the
          launching stream, max over ranks)
  e2e   : cells/s through the reference-facing C-ABI call (afq_submit/afq_wait) with pinned
          HOST buffers: H2D of the batch and D2H of the sparse result inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches : see DESIGN.md §Measurement

`--impl reference` times the reference's CPU algorithm (the oracle port, oracle/) on the host
cores of the box, on a bounded sample of the same workload.
"""
import argparse
import json
import os

# libafq keeps ~10 streams busy; see afq_more_hardware_queues() in afq_cuda.cu (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/sec quant (collated RAD→matrix) at 1/2/4/8 B200 vs CPU ref"
HBM_FALLBACK_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="afq", choices=["afq", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--cells", type=int, default=0, help="cells per GPU (default: the config's)")
    ap.add_argument("--resolution", default="")
    ap.add_argument("--e2e-batches", type=int, default=0,
                    help="host batches per e2e step (pipelined); 0 = 8 for cr-like / trivial, 4 for the graph-based and EM resolutions, "
                         "whose per-batch stage tails are longer (r2w: C4 61.3 ms at 4, 70.8 ms at 8; C2 36.7 vs 35.9)")
    ap.add_argument("--cpu-sample-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--others", default="C3,C4,C5", help="configurations measured after the default C2 headline (other_configs)")
    ap.add_argument("--no-others", action="store_true", help="measure the headline configuration only")
    ap.add_argument("--e2e-drain", action="store_true", help="e2e: collect every in-flight batch at the end of each step (no carry-over)")
    ap.add_argument("--e2e-offsets", action="store_true",
                    help="e2e: ship 4-byte rec_ref_offsets instead of the 1-byte rec_na8 alignment counts")
    ap.add_argument("--e2e-u32", action="store_true",
                    help="e2e: ship 4-byte UMIs / transcript ids instead of the 24-bit rec_umi24 / refs24 wire arrays")
    return ap.parse_args()


CONFIGS = {  # name -> (cells per GPU, resolution, description)
    "C1": (1000, "cr-like", "C1: 1k cells x 50k reads, 10x-v3, cr-like"),
    "C2": (100_000, "cr-like", "C2: 100k cells x ~200M reads per GPU, 10x-v3 16bp BC/12bp UMI, --resolution cr-like"),
    "C3": (125_000, "parsimony", "C3: 1M cells x 2B reads over 8 GPUs (125k cells/GPU), --resolution parsimony"),
    "C4": (125_000, "cr-like-em", "C4: 500k cells USA-mode over 4 GPUs (125k cells/GPU), --resolution cr-like-em"),
    "C5": (100_000, "parsimony-em", "C5: 100k cells, 40 reads/UMI, 5k genes, --resolution parsimony-em"),
}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        v = float(json.load(open(p))["hbm_gbs"])
        return v, "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                self.p.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes(batch, nnz):
    """every input byte read once, every output byte written once."""
    n_in = 8 * (batch.n_cells + 1) + 4 * batch.n_records + 4 * (batch.n_records + 1) + 4 * batch.n_refs_total
    n_out = 8 * nnz + 8 * (batch.n_cells + 1) + 17 * batch.n_cells
    return n_in + n_out


def build_hash():
    """sha256 (16 hex) over the CUDA sources libafq.so is built from: ties profiles/ncu_traffic.json to a build."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "alevin_fry_b200", "csrc", "*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def workload_config(synth, name, spec, cells, res, world, rank=0):
    """The `config` object of the JSON line — identical keys and values on both arms (the sizes come from the
    generator's size pass, which costs a fraction of a second and needs no GPU)."""
    n_rec, n_ref = synth.sizes(spec, rank * cells, cells)
    return {"workload": CONFIGS[name][2], "resolution": res, "cells_per_gpu": cells, "records_per_gpu": n_rec,
            "refs_per_gpu": n_ref, "parallelism": f"cell-sharded x{world}",
            "l2_policy": "inputs (%.1f GB/GPU) larger than L2 (126 MB); no flush needed" % ((8 * n_rec + 4 * n_ref) / 1e9),
            "seed": spec.seed}


def reference_leg(args, name, cells, res, steps, warmup, step_seconds):
    """The reference's CPU algorithm (oracle port) on all host cores, on a bounded sample of config `name`."""
    import oracle_lib
    from alevin_fry_b200 import QuantOpts
    import synth
    spec = synth.config_spec(name)
    cores = os.cpu_count() or 1
    t2g = synth.tid_to_gid(spec)
    opts = QuantOpts(resolution=res, usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows)
    probe = synth.generate(spec, 0, min(cells, 2048))
    oracle_lib.oracle_quant(opts, t2g, probe.slice_cells(0, min(64, probe.n_cells)), n_threads=cores)   # thread start-up
    t0 = time.perf_counter(); oracle_lib.oracle_quant(opts, t2g, probe, n_threads=cores); dt = time.perf_counter() - t0
    rate = probe.n_cells / dt
    n_sample = int(max(256, min(cells, rate * step_seconds)))
    sample = synth.generate(spec, 0, n_sample)
    for _ in range(warmup):
        oracle_lib.oracle_quant(opts, t2g, sample, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_lib.oracle_quant(opts, t2g, sample, n_threads=cores)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    v = n_sample / dt
    desc = f"first {n_sample} cells ({sample.n_records} records) of the workload per step, {steps} steps after {warmup} warm-up"
    # the CPU path at ONE worker as well (the reference's `-t 2` = 1 reader + 1 worker, BASELINE's "1 thread"): a ~2 s sample
    one = sample.slice_cells(0, int(max(64, min(n_sample, rate / max(cores, 1) * 2.0))))
    t0 = time.perf_counter(); oracle_lib.oracle_quant(opts, t2g, one, n_threads=1); dt1 = time.perf_counter() - t0
    return {
        "value": v, "unit": "cells/s", "ms_per_step": dt * 1e3,
        "config": workload_config(synth, name, spec, cells, res, args.gpus),
        "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": desc,
                         "one_worker": {"value": one.n_cells / dt1, "unit": "cells/s", "cores": 1, "sample": f"first {one.n_cells} cells, one pass"},
                         "note": "C++ restatement of the reference's Rust algorithm (reference not buildable here: no cargo/rustc)"},
        "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args, names, emit):
    """--impl reference: rank 0 alone runs; the headline config with the requested steps, the other configs briefly."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    total = args.steps + args.warmup
    legs = {}
    for i, name in enumerate(names):
        cells, res = CONFIGS[name][0], CONFIGS[name][1]
        if args.cells:
            cells = args.cells
        if args.resolution and i == 0:
            res = args.resolution
        if i == 0:
            legs[name] = reference_leg(args, name, cells, res, args.steps, args.warmup, min(8.0, 150.0 / max(total, 1)))
        else:
            legs[name] = reference_leg(args, name, cells, res, 2, 1, 5.0)
    hl = legs[names[0]]
    line = {"impl": "reference", "metric": METRIC, "value": hl["value"], "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": hl["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": hl["config"], "cpu_baseline": hl["cpu_baseline"], "e2e": hl["e2e"]}
    if len(names) > 1:
        line["other_configs"] = {n: legs[n] for n in names[1:]}
    emit(line)


def measure_config(args, name, cells, res, ctx):
    """One configuration on this rank's GPU: device-resident value, e2e through the host C-ABI, roofline of the
    per-cell resolve family, CPU baseline (rank 0, N=1). Returns the dict that becomes the line / an other_configs entry."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from alevin_fry_b200 import QuantOpts, Quantifier
    import synth
    from alevin_fry_b200.hostmem import PinnedPool
    world, rank, local, dev = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]
    steps, warmup = ctx["steps"], ctx["warmup"]
    spec = synth.config_spec(name)
    desc = CONFIGS[name][2]

    # ---- synthetic workload: this rank's shard of cells, generated into pinned host memory ----
    pool = PinnedPool()
    nthreads = max(1, (os.cpu_count() or 1) // max(1, min(world, 8)))
    nb = max(1, min(args.e2e_batches or (8 if res in ("cr-like", "trivial") else 4), cells))
    bounds = [cells * i // nb for i in range(nb + 1)]
    # each part: pinned arrays with part-relative offsets (what a host RAD parser would hand over)
    parts = [synth.generate(spec, rank * cells + bounds[i], bounds[i + 1] - bounds[i], n_threads=nthreads, alloc=pool.empty)
             for i in range(nb)]
    t2g = synth.tid_to_gid(spec)
    opts = QuantOpts(resolution=res, usa_mode=spec.usa_mode, num_gene_ids=spec.num_gene_ids, num_rows=spec.num_rows, device=local)
    q = Quantifier(opts, t2g)

    # device-resident copy of the whole per-GPU workload as ONE batch (offsets re-based on the device)
    def T(a, dt):
        return torch.from_numpy(a.view(dt)).to(dev)
    cro, umi, roff, refs, rbase, fbase = [], [], [], [], 0, 0
    for i, p in enumerate(parts):
        c = T(p.cell_rec_offsets, np.int64) + rbase
        r = T(p.rec_ref_offsets, np.int32).to(torch.int64) + fbase
        cro.append(c if i == nb - 1 else c[:-1])
        roff.append(r if i == nb - 1 else r[:-1])
        umi.append(T(p.rec_umi32, np.int32))
        refs.append(T(p.refs, np.int32))
        rbase += p.n_records
        fbase += p.n_refs_total
    if fbase >= 2 ** 32 - 16:
        raise SystemExit("per-GPU workload exceeds 2^32 alignments; lower --cells")
    roff64 = torch.cat(roff)
    roff32 = torch.where(roff64 >= 2 ** 31, roff64 - 2 ** 32, roff64).to(torch.int32)  # u32 bit pattern in an int32 tensor
    db = dict(cell_rec_offsets=torch.cat(cro), rec_umi32=torch.cat(umi), rec_ref_offsets=roff32, refs=torch.cat(refs))
    del cro, umi, roff, refs, roff64, roff32

    class batch:  # whole-workload sizes
        n_cells, n_records, n_refs_total = cells, rbase, fbase
    nc = batch.n_cells

    def out_set():
        return dict(row_ptr=torch.empty(nc + 1, dtype=torch.int64, device=dev),
                    col=torch.empty(batch.n_refs_total, dtype=torch.int32, device=dev),
                    val=torch.empty(batch.n_refs_total, dtype=torch.float32, device=dev),
                    sum_umi=torch.empty(nc, dtype=torch.float32, device=dev), max_umi=torch.empty(nc, dtype=torch.float32, device=dev),
                    num_expr=torch.empty(nc, dtype=torch.int32, device=dev), num_over_mean=torch.empty(nc, dtype=torch.int32, device=dev),
                    flags=torch.empty(nc, dtype=torch.uint8, device=dev))
    # N > 1: two output sets, so that the NCCL assembly of step i's matrix (on its own stream) overlaps step i+1's kernels
    outs = [out_set() for _ in range(2 if world > 1 else 1)]
    do = outs[0]
    asm_scratch = [{} for _ in outs]
    main = torch.cuda.current_stream()
    stream = main.cuda_stream
    comm = torch.cuda.Stream(priority=-1) if world > 1 else None      # (high priority, see init_process_group above)
    ev_done = [torch.cuda.Event() for _ in outs]       # step's kernels finished (main stream)
    ev_asm = [None for _ in outs]                      # step's matrix assembled (comm stream)
    pending = []                                       # output sets whose assembly has not been issued yet

    def assemble(j):
        # assembly of the ONE sparse matrix of the job (north_star: "an NCCL all-gather only for the final sparse matrix
        # assembly"): row lengths + the (col, val) payload of every rank, gathered over NVLink onto every rank
        from alevin_fry_b200 import shard
        with torch.cuda.stream(comm):
            comm.wait_event(ev_done[j])
            shard.assemble_csr(outs[j]["num_expr"], outs[j]["col"], outs[j]["val"], nc * world, scratch=asm_scratch[j], compact=False)
            ev_asm[j] = torch.cuda.Event()
            ev_asm[j].record(comm)

    step_no = [0]

    def step_device():
        j = step_no[0] % len(outs)
        step_no[0] += 1
        if ev_asm[j] is not None:
            main.wait_event(ev_asm[j])                 # the set is free once its previous matrix has been gathered
        q.quant_device(db, outs[j], stream)
        if world > 1:
            ev_done[j].record(main)
            # the previous step's assembly is issued AFTER this step's kernels are queued: its host-side waits (row lengths,
            # nnz) then overlap this step's compute, and so do its all-gathers
            while pending:
                assemble(pending.pop(0))
            pending.append(j)

    def drain_device():
        while pending:
            assemble(pending.pop(0))
        for e in ev_asm:
            if e is not None:
                main.wait_event(e)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------
    for _ in range(warmup):
        step_device()
    drain_device()
    nnz = q.device_finish(stream, outs[(step_no[0] - 1) % len(outs)]["row_ptr"])
    q.set_profiling(True)
    q.profile_reset()
    launches0 = q.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_device()
    drain_device()          # every step's matrix is assembled inside the timed region
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    launches_total = q.launch_count - launches0                 # kernels of libafq launched inside the timed region
    launches = launches_total // max(steps, 1)
    prof = q.profile()
    q.set_profiling(False)
    q.device_finish(stream)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- e2e through the host C-ABI: pinned host batches, H2D + kernels + D2H per step ---------
    use_na8 = not args.e2e_offsets
    if use_na8:  # the host keeps what a RAD record header stores (the alignment count, 1 byte) in pinned memory
        for p in parts:
            na = pool.empty(p.n_records, np.uint8)
            np.subtract(p.rec_ref_offsets[1:], p.rec_ref_offsets[:-1], out=na, casting="unsafe")
            p._na8 = na
    use_p24 = not args.e2e_u32   # 12-base UMIs and < 2^24 transcripts: 3 bytes each on the wire
    if use_p24:
        for p in parts:
            p.pack24(pool.empty)
    h2d = sum(p.cell_rec_offsets.nbytes + (p._na8.nbytes if use_na8 else p.rec_ref_offsets.nbytes) +
              (3 * (p.n_records + p.n_refs_total) if use_p24 else p.rec_umi32.nbytes + p.refs.nbytes) for p in parts)

    # A quant job is one stream of host batches: up to 3 are in flight (afq.h), and the pipeline is NOT drained between
    # steps — the batches still in flight when a step's last one is submitted are collected during the next step, the
    # last step's before the closing barrier. Every step's results are read back inside the timed region.
    tickets, acc = [], [0]

    def collect(t):
        n_c, n_z = q.wait(t, copy=False)
        acc[0] += 8 * (n_c + 1) + 17 * n_c + 8 * n_z

    def step_e2e():
        for p in parts:
            tickets.append(q.submit(p, use_na8, use_p24))
            if len(tickets) == 3:
                collect(tickets.pop(0))
        if args.e2e_drain:
            while tickets:
                collect(tickets.pop(0))
    for _ in range(warmup):
        step_e2e()
    while tickets:
        collect(tickets.pop(0))
    barrier()
    acc[0] = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    while tickets:
        collect(tickets.pop(0))
    barrier()
    d2h = acc[0] // steps
    e2e_s = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel family (per-cell resolve) -----------------------------
    fam = {k: v for k, v in prof.items() if k.startswith(("k_resolve", "k_gene_eqc", "k_pug_smem", "k_pug_build", "k_pug_cover", "k_pug_count", "k_pug_back", "k_back_bin", "k_em_cells", "k_em_bin"))}
    fam_ms = sum(v[0] for v in fam.values()) / steps
    region = prof.get("resolve_region(wall)")
    pug_region = prof.get("pug_region(wall)")
    if region:  # arena kernels overlap on lanes: their device time is the wall time of the region(s)
        fam_ms = region[0] / steps
        if pug_region:   # the k_pug_smem variants overlap on lanes too; k_gene_eqc (handed-back cells) runs behind them
            fam_ms += (pug_region[0] + sum(v[0] for k, v in fam.items() if k.startswith(("k_gene_eqc", "k_pug_count", "k_back_bin", "k_em_bin")))) / steps
            for reg in ("cover_region(wall)", "back_region(wall)", "em_region(wall)"):   # split path: the flat cover / back-end kernels overlap on lanes as well
                if prof.get(reg):
                    fam_ms += prof[reg][0] / steps
        else:
            fam_ms += sum(v[0] for k, v in fam.items() if k.startswith(("k_gene_eqc", "k_pug_smem"))) / steps
    fam_launches = sum(v[1] for v in fam.values()) // max(steps, 1)
    abytes = algorithmic_bytes(batch, nnz)
    peak, peak_src = peak_hbm()
    achieved = abytes / (fam_ms * 1e-3) / 1e9 if fam_ms > 0 else 0.0
    traffic, traffic_note, warp_inst = None, "no capture for this configuration", None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj.get("build") == build_hash():
            traffic = tj.get(name)
            traffic_note = tj.get("source")
            warp_inst = (tj.get("warp_instructions") or {}).get(name)
        else:
            traffic_note = "profiles/ncu_traffic.json was captured on another build of csrc/ (%s): not reported" % tj.get("build")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "per-cell resolve family (k_resolve_smem<*>/k_resolve_large/k_pug_smem<*>/k_pug_build<*>+k_pug_cover*+k_pug_count/k_gene_eqc), %d launches/step" % fam_launches,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                "algorithmic_bytes_per_step": abytes, "kernel_ms_per_step": fam_ms, "peak_source": peak_src, "nnz_per_gpu": nnz,
                "per_kernel_ms": {k: v[0] / steps for k, v in prof.items()}}
    # the family is instruction-issue / latency bound, not HBM bound: issue-slot utilisation beside the HBM fraction.
    # warp instructions per step from the same ncu pass as `traffic` (same build), peak = SMs x 4 schedulers x SM clock
    if warp_inst and traffic and fam_ms > 0:
        scale = cells / float(tj.get("cells_per_step", {}).get(name, cells))      # (the capture's step may be a smaller cell count)
        sm_hz = ((clocks or {}).get("sm_mhz") or 1965) * 1e6
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        roofline["issue"] = {"warp_instructions_per_step": warp_inst * scale, "achieved_ginst_s": warp_inst * scale / (fam_ms * 1e-3) / 1e9,
                             "peak_ginst_s": n_sm * 4 * sm_hz / 1e9, "frac": warp_inst * scale / (fam_ms * 1e-3) / (n_sm * 4 * sm_hz)}

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample of the same workload --
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib
        cores = os.cpu_count() or 1
        budget = ctx["cpu_seconds"]
        probe = parts[0].slice_cells(0, min(parts[0].n_cells, 512))
        t0 = time.perf_counter(); oracle_lib.oracle_quant(opts, t2g, probe, n_threads=cores); dt = time.perf_counter() - t0
        n_sample = int(max(256, min(parts[0].n_cells, probe.n_cells / dt * budget)))
        # the 512-cell probe is dominated by thread start-up and under-estimates the rate: re-size once so that
        # the timed sample is >= ~20 CPU-seconds of work (and >= 1.5 s of wall time), bounded by the workload
        for attempt in range(2):
            sample = parts[0].slice_cells(0, n_sample)
            t0 = time.perf_counter(); oracle_lib.oracle_quant(opts, t2g, sample, n_threads=cores); dt = time.perf_counter() - t0
            want_wall = max(1.5, 20.0 / cores)
            if attempt == 1 or dt >= 0.7 * want_wall or n_sample >= parts[0].n_cells:
                break
            n_sample = int(min(parts[0].n_cells, max(n_sample + 1, n_sample * want_wall / max(dt, 1e-3))))
        cpu = {"value": n_sample / dt, "unit": "cells/s", "cores": cores, "kind": "port",
               "sample": f"first {n_sample} cells ({sample.n_records} records) of the workload, {dt:.1f} s of wall time on {cores} threads",
               "note": "C++ restatement of the reference's Rust algorithm (reference not buildable here: no cargo/rustc)"}

    # ---- tie census (parsimony family, rank 0, N=1): how much of the result could depend on the cover's visiting order,
    # the one place where the reference follows hash-iteration order -------------------------------------
    census = None
    if cpu is not None and res.startswith("parsimony"):
        import oracle_lib
        cs = parts[0].slice_cells(0, min(parts[0].n_cells, 2000))
        census = oracle_lib.tie_census(opts, t2g, cs, n_threads=os.cpu_count() or 1)
        census["sample"] = f"first {cs.n_cells} cells of the workload"

    total_cells = nc * world
    cfg = workload_config(synth, name, spec, cells, res, world, rank)
    out = {
        "value": total_cells / (ms * 1e-3), "unit": "cells/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
        "config": cfg, "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": total_cells / e2e_s, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "host_batches_per_step": nb, "pipeline": "drained every step" if args.e2e_drain else "3 batches in flight, carried across steps, drained before the closing barrier",
                "input_encoding": ("rec_umi24 + " if use_p24 else "rec_umi32 + ") + ("rec_na8 + " if use_na8 else "rec_ref_offsets + ") +
                                  ("refs24" if use_p24 else "refs (u32)")},
        "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches), "clocks": clocks,
    }
    if census is not None:
        out["tie_census"] = census
    if world > 1:
        out["assembly"] = ("NCCL all-gather of row lengths + (col, val) onto every rank, inside the timed step; issued on a second stream so that "
                           "step i's gathers overlap step i+1's kernels (two output sets), all drained before the closing event")
    q.close()
    del db, do, outs, asm_scratch, parts
    pool.close()
    torch.cuda.empty_cache()
    return out


def main():
    args = parse_args()
    # the contract is ONE JSON line on stdout: route everything else (NCCL banners, library chatter)
    # to stderr until the line is printed
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)
    # the headline configuration first; a default run (C2) also measures the parsimony / EM configurations of
    # BASELINE.json (configs[2..4]) and reports them under "other_configs" on both arms
    names = [args.config]
    if args.config == "C2" and not args.no_others and not args.cells and not args.resolution:
        names += [n for n in args.others.split(",") if n and n in CONFIGS and n != "C2"]
    if args.impl == "reference":
        return run_reference(args, names, emit)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the afq product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the per-cell kernels are persistent and fill every SM: NCCL's kernels (and the small copy kernels of the assembly) run
        # on HIGH-PRIORITY streams so that they get the first CTA slot a finishing kernel frees instead of queueing behind the
        # whole step (r2aa: without, the gathers of step i ran behind step i+1's kernels, 9.18 ms at N=2 vs 8.2 ms of compute)
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    ctx = {"world": world, "rank": rank, "local": local, "dev": dev, "steps": args.steps, "warmup": max(args.warmup, 3),
           "cpu_seconds": args.cpu_sample_seconds}
    legs = {}
    for i, name in enumerate(names):
        cells, res = CONFIGS[name][0], CONFIGS[name][1]
        if args.cells:
            cells = args.cells
        if args.resolution:
            res = args.resolution
        if i > 0:   # the other configurations: 3 warm-up + 3 timed steps, a shorter CPU sample
            ctx = dict(ctx, steps=min(args.steps, 3), warmup=3, cpu_seconds=min(args.cpu_sample_seconds, 6.0))
        legs[name] = measure_config(args, name, cells, res, ctx)
    if rank == 0:
        hl = legs[names[0]]
        line = {"metric": METRIC, "value": hl["value"], "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": hl["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": hl["config"], "roofline": hl["roofline"],
                "cpu_baseline": hl["cpu_baseline"], "e2e": hl["e2e"], "gpu_launches": hl["gpu_launches"], "gpu_launches_per_step": hl["gpu_launches_per_step"],
                "clocks": hl["clocks"]}
        if hl.get("assembly"):
            line["assembly"] = hl["assembly"]
        if len(names) > 1:
            line["other_configs"] = {n: legs[n] for n in names[1:]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
