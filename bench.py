#!/usr/bin/env python
"""Benchmark of the Fitch/Sankoff construction pass (BASELINE.json metric: node x column updates / s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--algo fitch|sankoff] [--impl reference]

A "step" is one pass of the hot path (forward + backward + ordered mutation compaction) over one batch of
synthetic columns. N=1 runs BASELINE.json configs[1] (synthetic SARS-CoV-2-like MSA, 20k leaves x 30k columns,
random binary tree, Fitch). For N>1 (launched by torch.distributed.run, one rank per GPU) every rank owns one such
30k-column range of an N x 30k-column alignment on the replicated tree -- per-GPU work is fixed, scaling "weak" --
and each step ends with the column-range gather of the per-rank mutation lists to rank 0 over NCCL.
Prints ONE JSON line (rank 0).  --impl reference times the reference's own CPU implementation instead.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="sars20k")
    ap.add_argument("--algo", default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cols", type=int, default=0, help="override the column count (debug)")
    ap.add_argument("--leaves", type=int, default=0, help="override the leaf count (debug)")
    ap.add_argument("--chunk-nodes", type=int, default=0)
    ap.add_argument("--col-groups", type=int, default=0)
    ap.add_argument("--reserve-sms", type=int, default=8,
                    help="N > 1: SMs the pass kernels leave free so that the NCCL gather of step i can overlap pass i+1 (0 = serialise)")
    ap.add_argument("--e2e-contexts", type=int, default=1,
                    help="N = 1: contexts (host threads) the end-to-end measurement streams its batches through (1 = one call at a time)")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: split the configuration's columns over the ranks (strong scaling) instead of one configuration per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload(args):
    from panman_b200 import synth

    cfg = dict(synth.CONFIGS[args.config])
    if args.cols:
        cfg["n_cols"] = args.cols
    if args.leaves:
        cfg["n_leaves"] = args.leaves
    algo = args.algo or cfg["algos"][0]
    return cfg, algo


# ----------------------------------------------------------------------------- clocks (sampled during the timed region)
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, k), int)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bit and (r & bit) and name not in ("nvmlClocksThrottleReasonNone", "nvmlClocksThrottleReasonAll"):
                        self.reasons.add(name.replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- reference / oracle CPU timing
def cpu_reference_run(tree, codes_sample, parent_code_sample, algo, seconds, threads):
    """Times the reference's own CPU implementation of the path on a bounded column sample: oracle/_ref (the
    verbatim fitchSankoff.cpp + restated string-keyed drivers) when it was built, else the oracle port."""
    from oracle.oracle import CHAR_OF, PortOracle, RefOracle, have_ref

    n_cols = codes_sample.shape[1]
    algo_i = 0 if algo == "fitch" else 1
    names = tree.names()
    if have_ref():
        ref = RefOracle()

        class T:  # the fields RefOracle.tree needs
            pass

        t = T()
        t.n_nodes, t.names, t.parent, t.child_off, t.child_idx = tree.n_nodes, names, tree.parent, tree.child_off, tree.child_idx
        h = ref.tree(t)
        leaf_names = [names[v] for v in tree.leaves]
        rows = CHAR_OF[codes_sample]
        cons = CHAR_OF[parent_code_sample]

        def run(nc):
            t0 = time.perf_counter()
            ref.msa_run(h, t, algo_i, leaf_names, [bytes(r[:nc]) for r in rows], bytes(cons[:nc]), "", n_threads=threads)
            return time.perf_counter() - t0

        kind = "reference"
    else:
        port = PortOracle()

        def run(nc):
            t0 = time.perf_counter()
            port.run(tree, algo_i, codes_sample[:, :nc], parent_code_sample[:nc], n_threads=threads)
            return time.perf_counter() - t0

        kind = "port"
    # grow the sample until one run takes about `seconds` (a short probe overweights the fixed costs and undershoots)
    nc = min(n_cols, max(threads, 16))
    dt = run(nc)
    while dt < 0.5 * seconds and nc < n_cols:
        nc = int(min(n_cols, max(2 * nc, 0.9 * nc * seconds / max(dt, 1e-6))))
        dt = run(nc)
    return dict(value=tree.n_nodes * nc / dt, unit="node*col/s", cores=threads, kind=kind,
                sample=f"first {nc} columns of the workload ({tree.n_nodes} nodes), {dt:.1f} s, "
                       f"{'string-keyed reference drivers' if kind == 'reference' else 'array port'}, {threads} threads")


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- main arms
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from panman_b200 import synth

    cfg, algo = workload(args)
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    sample_cols = int(min(cfg["n_cols"], 8192, max(512, 300_000_000 // cfg["n_leaves"])))
    codes4, pc = synth.simulate_msa(tree, 0, sample_cols, synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"]))
    codes = synth.unpack_nibbles(codes4, sample_cols).numpy()
    threads = host_threads()
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_run(tree, codes, pc.numpy(), algo, per_step, threads)
        if i >= args.warmup:
            vals.append(last["value"])
    value = float(np.mean(vals))
    last["value"] = value
    line = {
        "impl": "reference", "metric": "fitch_sankoff_node_column_updates_per_sec", "value": value, "unit": "node*col/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tree.n_nodes * cfg["n_cols"] / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16 sets / int32 costs (CPU)", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['n_leaves']} leaves x {cfg['n_cols']} columns, {cfg['kind']} tree, {algo}; "
                               "each step = bounded column sample, rate extrapolated (columns are independent)"},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": "node*col/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def run_b200_arm(args):
    import torch

    import panman_b200 as pb
    from panman_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libpanman_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "4")  # the gather moves a few MB: few channels = few SMs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, algo = workload(args)
    algo_i = pb.ALGO_FITCH if algo == "fitch" else pb.ALGO_SANKOFF
    C = cfg["n_cols"]
    c0 = rank * C  # weak scaling: rank r owns columns [r*C, (r+1)*C) of a world*C-column alignment
    if args.strong and world > 1:
        # strong scaling (BASELINE.json configs[3]: ONE alignment column-sharded over the GPUs): tile-aligned contiguous ranges
        tiles = (cfg["n_cols"] + 1023) // 1024
        a, b = tiles * rank // world, tiles * (rank + 1) // world
        c0, C = a * 1024, min(cfg["n_cols"], b * 1024) - a * 1024
    tree = synth.make_tree(cfg["n_leaves"], cfg["seed"], cfg["kind"])
    spec = synth.MsaSpec(cfg["seed"], cfg["p_sub"], cfg["f_gap"], cfg["p_N"])
    dev = torch.device("cuda", local)
    codes4, pc = synth.simulate_msa(tree, c0, c0 + C, spec, device=dev)
    ro = None
    if algo == "sankoff":  # Sankoff runs with --reference = leaf 0
        ro = (synth.unpack_nibbles(codes4[:1], C)[0]).to(torch.int8).contiguous()
    torch.cuda.synchronize()

    ctx = pb.Context(local)
    if args.chunk_nodes:
        ctx.set_option("chunk_nodes", args.chunk_nodes)
    if args.col_groups:
        ctx.set_option("col_groups", args.col_groups)
    overlap_gather = world > 1 and args.reserve_sms > 0
    if overlap_gather:
        ctx.set_option("reserve_sms", args.reserve_sms)
    ctx.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
    ctx.upload(C, tree.n_leaves, codes4, codes4.shape[1], pc, ro, None, None, c0)

    N = tree.n_nodes

    # ---- N > 1: every step ends with the column-range gather of the per-rank lists to rank 0 and their merge there.
    # A rank packs its result into one device buffer (pmb_pack_result) -> ONE NCCL gather -> pmb_merge_packed on rank 0.
    # The gather/merge of step i runs on a side stream and overlaps the passes of step i+1 (double-buffered); the
    # timed region ends only after the last merge has finished.
    shard = None
    if world > 1:
        ctx.run_resident(algo_i)
        nmax = torch.tensor([ctx.result_device().n_mut], dtype=torch.int64, device=dev)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
        cap = int(nmax.item() * 5 // 4) + 4096  # same on every rank; identical steps => stable
        pbytes = ctx.packed_bytes(cap)
        shard = dict(cap=cap, bytes=pbytes, send=[torch.empty(pbytes, dtype=torch.uint8, device=dev) for _ in range(2)],
                     recv=[torch.empty(world * pbytes, dtype=torch.uint8, device=dev) for _ in range(2)] if rank == 0 else None,
                     comm=torch.cuda.Stream(device=dev), lib=torch.cuda.ExternalStream(ctx.stream_handle(), device=dev), i=0,
                     merged=None, done=[None, None], events=[torch.cuda.Event(), torch.cuda.Event()],
                     merge_stream=torch.cuda.Stream(device=dev), mdone=[None, None], mevents=[torch.cuda.Event(), torch.cuda.Event()])
        # receive views built once: the step loop is host-bound at N > 1, every Python object per step counts
        shard["dst"] = [[shard["recv"][k][r * pbytes:(r + 1) * pbytes] for r in range(world)] for k in range(2)] if rank == 0 else [None, None]

    def gather_lists():
        k = shard["i"] & 1
        shard["i"] += 1
        comm, lib = shard["comm"], shard["lib"]
        if shard["done"][k] is not None:
            lib.wait_event(shard["done"][k])                # the gather that last used this buffer pair has finished
        ctx.pack_result(shard["send"][k], shard["cap"])     # on the library's stream, right behind the pass
        comm.wait_stream(lib)
        with torch.cuda.stream(comm):
            if rank == 0 and shard["mdone"][k] is not None:
                comm.wait_event(shard["mdone"][k])          # the merge that last read this receive buffer has finished
            dist.gather(shard["send"][k], shard["dst"][k], dst=0)  # enqueued on comm; the host does not block
            done = shard["events"][k]
            done.record(comm)
        if rank == 0:
            # the merge runs on its own stream: the gather of step i+1 overlaps the merge of step i (at 8 ranks the two in a
            # row on one stream took longer than a pass and throttled the whole pipeline)
            ms = shard["merge_stream"]
            ms.wait_event(done)
            shard["merged"] = ctx.merge_packed(world, shard["recv"][k], shard["cap"], stream=ms.cuda_stream)
            shard["mevents"][k].record(ms)
            shard["mdone"][k] = shard["mevents"][k]
        shard["done"][k] = done
        # The persistent pass kernels fill every SM they are given, and an NCCL kernel that has to squeeze in beside them
        # (on both ranks at once) stalls far longer than it runs. Either the pass kernels leave a few SMs free
        # (--reserve-sms) and the gather of step i overlaps pass i+1, or the next pass waits for the collective.
        # Rank 0's merge kernels are small and overlap the next pass in both cases.
        if not overlap_gather:
            lib.wait_event(done)

    lib_stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)

    def step():
        """One pass, enqueued asynchronously (pmb_run_resident_async): passes run back to back on the library's stream
        and, for N > 1, the gather of a finished pass overlaps the next one. Nothing is skipped: pmb_wait at the end
        checks the status of every pass."""
        ctx.run_resident_async(algo_i)
        if world > 1:
            gather_lists()

    def barrier():
        ctx.wait()
        if world > 1:
            shard["comm"].synchronize()
            shard["merge_stream"].synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(lib_stream)
    for _ in range(args.steps):
        step()
    ev1.record(lib_stream)
    barrier()
    elapsed = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    tot = ev0.elapsed_time(ev1)  # device time of the K passes on the stream they were launched on
    launches = ctx.timings().n_launches * args.steps
    # phase split of one pass: a few synchronous passes after the timed region (same kernels, same inputs)
    fwd = bwd = cmp_ = 0.0
    for _ in range(3):
        t = ctx.run_resident(algo_i)
        fwd += t.forward_ms / 3
        bwd += t.backward_ms / 3
        cmp_ += t.compact_ms / 3
    n_mut = ctx.download(copy=False).n_mut
    alg_bytes = ctx.algorithmic_bytes(algo_i)
    el = torch.tensor([elapsed, tot / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed, dev_total = float(el[0]), float(el[1])
    K = args.steps
    units = N * (cfg["n_cols"] if (args.strong and world > 1) else C * world)
    value = units * K / elapsed

    # ---- end to end through the reference-facing C-ABI call with HOST buffers (H2D and D2H inside the timed region)
    h_codes = torch.empty(codes4.shape, dtype=torch.uint8, pin_memory=True).copy_(codes4)
    h_pc = torch.empty(pc.shape, dtype=torch.uint8, pin_memory=True).copy_(pc)
    h_ro = None if ro is None else torch.empty(ro.shape, dtype=torch.int8, pin_memory=True).copy_(ro)
    torch.cuda.synchronize()
    e2e_steps = max(2, min(K, 6))
    res = ctx.run_nuc(algo_i, C, tree.n_leaves, h_codes, h_codes.shape[1], h_pc, h_ro, None, None, c0, 0, copy=False)
    barrier()
    e2e_api = "pmb_run_nuc with pinned host buffers"
    if world == 1 and args.e2e_contexts > 1:
        # Batches are independent, so a caller streams them through two contexts on two host threads (one context per
        # thread is the library's threading model): the upload of one batch overlaps the pass and download of the other
        # and the PCIe link, the bound of this path, stays busy. Every call still moves its own inputs and results.
        ctxs = [ctx]
        for _ in range(args.e2e_contexts - 1):
            c2 = pb.Context(local)
            c2.set_tree(tree.n_nodes, tree.root, tree.child_off, tree.child_idx, tree.leaf_row)
            c2.run_nuc(algo_i, C, tree.n_leaves, h_codes, h_codes.shape[1], h_pc, h_ro, None, None, c0, 0, copy=False)
            ctxs.append(c2)
        torch.cuda.synchronize()
        per = max(2, e2e_steps // len(ctxs) + 1)
        results = [None] * len(ctxs)

        def worker(i):
            torch.cuda.set_device(local)
            for _ in range(per):
                results[i] = ctxs[i].run_nuc(algo_i, C, tree.n_leaves, h_codes, h_codes.shape[1], h_pc, h_ro, None, None, c0, 0, copy=False)

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(ctxs))]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        e2e_t = time.perf_counter() - t0
        e2e_steps = per * len(ctxs)
        res = results[0]
        e2e_api = f"pmb_run_nuc with pinned host buffers, {len(ctxs)} contexts on {len(ctxs)} host threads (batches streamed)"
        for c2 in ctxs[1:]:
            c2.close()
    else:
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = ctx.run_nuc(algo_i, C, tree.n_leaves, h_codes, h_codes.shape[1], h_pc, h_ro, None, None, c0, 0, copy=False)
            if world > 1:
                gather_lists()
        barrier()
        e2e_t = time.perf_counter() - t0
    e2 = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    e2e_value = units * e2e_steps / float(e2[0])
    h2d = int(h_codes.numel() + h_pc.numel() + (0 if h_ro is None else h_ro.numel()))
    d2h = int((N + 1) * 8 + res.n_mut * 5)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    pass_ms = 1e3 * dev_total / K  # CUDA events around the K timed passes on the library's stream
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.config}:{algo}")
    except Exception:
        pass
    line = {
        "metric": "fitch_sankoff_node_column_updates_per_sec", "value": value, "unit": "node*col/s", "n_gpus": world,
        "steps": K, "warmup": max(3, args.warmup), "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True,
        "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
        "dtype": "u16 Fitch sets as 16 bit-planes" if algo == "fitch" else "2-bit Sankoff excess as 32 bit-planes", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['n_leaves']} leaves ({N} nodes) x {C} columns per GPU, {cfg['kind']} tree, "
                               f"{algo}, seed {cfg['seed']}; " + (f"{cfg['n_cols']} columns split over the ranks" if (args.strong and world > 1) else f"rank r owns columns [r*{C},(r+1)*{C})"),
                   "l2": "inputs larger than L2: leaf planes + set matrix = "
                         f"{(tree.n_leaves * 0.5 + (N - tree.n_leaves) * (2 if algo == 'fitch' else 4)) * C / 1e6:.0f} MB per pass",
                   "parallelism": (f"column ranges x{world}, tree replicated, NCCL gather of mutation lists"
                                   + (f" overlapping the next pass ({args.reserve_sms} SMs left free)" if overlap_gather else "")) if world > 1 else "single GPU",
                   "n_mut_rank0": int(n_mut)},
        "device_ms_per_step": 1e3 * dev_total / K,
        "phases_ms": {"forward": fwd, "backward": bwd, "compact": cmp_, "note": "3 synchronous passes after the timed region"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "node*col/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": e2e_api},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "one pass = persistent forward kernel + persistent backward kernel + compaction; "
                               "duration = CUDA events around the timed passes on the library's stream / steps",
                     "algorithmic_bytes_per_pass": int(alg_bytes)},
    }
    if not args.no_cpu_baseline:
        sample_cols = int(min(C, 16384, max(512, 300_000_000 // tree.n_leaves)))  # bounded by host memory; ~cpu_seconds of work
        codes = synth.unpack_nibbles(codes4[:, :(sample_cols + 1) // 2], sample_cols).cpu().numpy()
        line["cpu_baseline"] = cpu_reference_run(tree, codes, pc[:sample_cols].cpu().numpy(), algo, args.cpu_seconds, host_threads())
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
