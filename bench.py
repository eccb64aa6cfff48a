#!/usr/bin/env python
"""bench.py — Gibbs sweep throughput of the mmseq hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload "C2-perfragment"): BASELINE.json configs[1] — a synthetic
Ensembl-sized sample, 180k transcripts / 30M paired fragments per GPU, one CSR row
per fragment (k == 1), rows grouped by hit class by the loader.  A "step" is
SWEEPS_PER_STEP (= 16, the reference's trace stride, src/mmseq.cpp:192) full Gibbs
sweeps: allocation of every hit class (k_alloc), [NCCL all-reduce of the count
vector when N > 1], Gamma update of every transcript (k_gamma), with one trace
capture per step.  Weak scaling: every rank holds its own 30M-fragment shard of the
same transcriptome; the aggregate metric is hit-class allocations/s.

Printed by rank 0: ONE JSON line (see README / DESIGN.md for the keys).
`--impl reference` times the reference's own algorithm and data flow on the host
cores (oracle/: MT19937 per OpenMP thread, GSL-style samplers, dense per-thread
partials; the reference itself needs Boost + GSL and cannot be built here).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SWEEPS_PER_STEP = 16
SEED = 1234
SYNTH_SEED = 20260101 + 2
T_C2 = 180_000
N_C2 = 30_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--transcripts", type=int, default=T_C2)
    ap.add_argument("--fragments", type=int, default=N_C2, help="fragments per GPU")
    ap.add_argument("--layout", default="perfragment", choices=["perfragment", "perfragment_byclass", "perfragment_unsorted", "collapsed"])
    ap.add_argument("--first-appearance-columns", action="store_true", help="number transcripts as the reference does (src/mmseq.cpp:403) instead of in header order")
    ap.add_argument("--weights", action="store_true", help="fp32 per-hit weights (config 4's extension)")
    ap.add_argument("--haplo", action="store_true", help="config 3: haplotype-specific transcriptome (every transcript as _A/_B copies: 360k haplo-transcripts, deep multi-mapping)")
    ap.add_argument("--transposed", action="store_true", help="materialise X + atomic-free transposed reduction")
    ap.add_argument("--cpu-sweeps", type=int, default=4, help="sweeps of the CPU baseline sample")
    ap.add_argument("--nccl-only", action="store_true", help="N > 1: exchange counts with ncclAllReduce instead of the fused peer-memory kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-collapsed", action="store_true", help="skip the extra line on the collapsed (reference-semantics) layout of the same sample")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_workload(args, rank, world):
    """This rank's shard: its own fragments of the shared transcriptome, columns = header indices
    when world > 1 (one column space across shards)."""
    from mmseq_b200 import hostlib, synth
    t0 = time.time()
    s = synth.Synth(SYNTH_SEED + (1 if args.haplo else 0), args.transcripts * (2 if args.haplo else 1), args.fragments, haplo=args.haplo, weights=args.weights, frag_seed=rank)
    t1 = time.time()
    layout = {"perfragment": hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH, "perfragment_byclass": hostlib.LAYOUT_PER_FRAGMENT_SORTED,
              "perfragment_unsorted": hostlib.LAYOUT_PER_FRAGMENT,
              "collapsed": hostlib.LAYOUT_COLLAPSED}[args.layout]
    if world > 1:
        layout |= hostlib.LAYOUT_IDENTITY_COLUMNS
    elif not args.first_appearance_columns:
        layout |= hostlib.LAYOUT_HEADER_ORDER_COLUMNS
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, frag_w=s.frag_w if args.weights else None, layout=layout)
    t2 = time.time()
    # l[t] = efflen * N_total / 1e9 over the whole (all-rank) sample, src/mmseq.cpp:603
    n_total = args.fragments * world
    length = s.efflen[h.col2hdr] * n_total / 1e9
    return s, h, length, dict(gen_s=round(t1 - t0, 2), load_s=round(t2 - t1, 2))


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.sm = []
        self.reasons = set()
        self.max_mhz = None
        self.ok = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            dev = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(dev, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.ok = True
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM))
                r = get(dev)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.005)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def result(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def algorithmic_bytes(h, n, weights, transposed, stride, layout, cls=None):
    """Compulsory HBM bytes of the implemented data flow ('Roofline'); every array
    once, L2-resident mu gathers and count reductions not counted.
      segment kernel (by-length k == 1 shards): 4 B column (+4 B weight) per CSR entry of the
        classes with >= 2 members; no row pointers, singleton classes are not visited;
      row-pointer kernels: + 8 B row pointer per class, all entries (+4 B k when present);
      transposed variant: + X written, + permutation and X read back (12 B per entry)."""
    m, nnz = h.m, h.nnz
    d = np.diff(h.row_ptr)
    per_entry = 4 + (4 if weights else 0)
    if cls and cls["in_use"] and not transposed:
        # class plan (mmq_cls.cu): packed columns (4 B) + draws/slot number (2 B) + class id (4 B) per slot of the small
        # set; row pointer, k, class id (8 + 4 + 8 B) per class and 4 B per entry of the sub-CSR left to k_alloc
        alloc = 4 * cls["packed_slots"] + 6 * cls["class_slots"] + 20 * cls["rest_classes"] + 4 * cls["rest_nnz"]
    elif layout == "perfragment" and h.k is None and not transposed:
        alloc = per_entry * int(d[d >= 2].sum())
    else:
        alloc = 8 * (m + 1) + per_entry * nnz + (4 * m if h.k is not None else 0)
    if transposed:
        alloc += 4 * nnz
    reduce_ = (8 * nnz + 8 * (n + 1) + 4 * n) if transposed else 0
    gamma = n * (4 + 4 + 8 + 8) + (8 * n) // stride
    return alloc, alloc + reduce_ + gamma


def run_reference(args, rank, world):
    """The reference's algorithm on the host cores (oracle/ port), same workload and metric."""
    if rank != 0:
        return
    from oracle import oracle as orc
    s, h, length, prep = make_workload(args, 0, 1)
    P = orc.Problem(h.row_ptr, h.col, h.k, length)
    mu, _, _ = P.init_mu()
    threads = orc.max_threads()
    mu, _, _ = P.gibbs_gsl(mu, SEED, max(1, args.warmup), threads=threads)  # warm-up sweeps
    secs = []
    for _ in range(args.steps):
        mu, _, sec = P.gibbs_gsl(mu, SEED, 1, threads=threads)  # one step of the reference arm = ONE sweep
        secs.append(sec)
    tot = float(np.sum(secs))
    sweeps_per_s = args.steps / tot
    value = sweeps_per_s * h.m
    line = {
        "impl": "reference", "metric": "gibbs_hit_class_allocations_per_s", "value": value, "unit": "allocations/s",
        "sweeps_per_s": sweeps_per_s, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, h, 1, sweeps_per_step=1),
        "cpu_baseline": {"value": value, "unit": "allocations/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full sweeps over all {h.m} classes of one shard (1 sweep per step)"},
        "e2e": {"value": value, "unit": "allocations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of src/mmseq.cpp:851-918 (MT19937 per OpenMP thread, GSL-style samplers, dense per-thread "
                "partials); the reference's own main() builds here only against the stand-in Boost/GSL of oracle/shim and needs minutes per sweep at this size (DESIGN.md section 9)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, h, world, sweeps_per_step=SWEEPS_PER_STEP):
    return {
        "workload": ("C3-haplotype-" if getattr(args, "haplo", False) else "C2-") + args.layout + ("-weighted" if args.weights else ""),
        "transcripts": args.transcripts, "fragments_per_gpu": args.fragments, "fragments_total": args.fragments * world,
        "n_columns": int(h.n), "classes_per_gpu": int(h.m), "nnz_per_gpu": int(h.nnz), "distinct_classes_per_gpu": int(h.n_classes),
        "sweeps_per_step": sweeps_per_step, "trace_stride": SWEEPS_PER_STEP, "seed": SEED,
        "count_path": "transposed" if args.transposed else "fused_reduction",
        "count_exchange": ("none" if world == 1 else ("nccl_allreduce" if getattr(args, "nccl_only", False) else "fused_p2p_gamma")),
        "l2": "inputs_exceed_l2" if (4 * h.nnz + 8 * h.m) > 200e6 else "inputs_fit_l2_flush_between_steps",
    }


def collapsed_line(args, s, dev, stream, length_full=None, steps=5, warmup=3):
    """Sweeps/s of the collapsed layout of the same synthetic sample (device-timed, CUDA events)."""
    import torch
    from mmseq_b200 import capi, hostlib
    h = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=hostlib.LAYOUT_COLLAPSED | hostlib.LAYOUT_HEADER_ORDER_COLUMNS)
    length = s.efflen[h.col2hdr] * args.fragments / 1e9
    rp_s, col_s, k_s, class_id = hostlib.sort_classes_by_cost(h)
    H = capi.Handle(rp_s, col_s, k_s, length, class_id=class_id, device=dev.index)
    H.set_stream(stream.cuda_stream)
    H.init_mu()
    cls = H.cls_stats()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # the packed classes fit the L2: flush between steps
    sweep = 0
    S = SWEEPS_PER_STEP

    def step(flags):
        nonlocal sweep
        with torch.cuda.stream(stream):
            flush.zero_()
        H.gibbs(SEED, sweep, S, stride=S, trace_len=steps + warmup + 1, flags=flags)
        sweep += S

    for _ in range(warmup):
        step(capi.MMQ_GIBBS_DEFAULT)
    H.synchronize()
    for _ in range(steps):
        step(capi.MMQ_GIBBS_TIME_KERNELS)
    alloc_ms, alloc_n, gamma_ms, gamma_n = H.kernel_times()
    H.close()
    per_sweep = alloc_ms / max(alloc_n, 1) + gamma_ms / max(gamma_n, 1)
    b = 4 * cls["packed_slots"] + 6 * cls["class_slots"] + 20 * cls["rest_classes"] + 4 * cls["rest_nnz"]
    return {"workload": "C2-collapsed", "classes": int(h.m), "nnz": int(h.nnz), "kernel": "k_alloc_cls",
            "alloc_avg_launch_ms": alloc_ms / max(alloc_n, 1), "gamma_avg_launch_ms": gamma_ms / max(gamma_n, 1),
            "sweeps_per_s_kernels": 1000.0 / per_sweep, "fragments_per_s_kernels": 1000.0 / per_sweep * args.fragments,
            "algorithmic_bytes_per_launch": int(b), "class_plan": cls,
            "note": "kernel time per sweep (allocation + Gamma, CUDA events), launch gaps not included; python bench.py --layout collapsed gives the full line"}


def main():
    args = parse()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mmseq_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the mmseq hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s, h, length, prep = make_workload(args, rank, world)
    n = h.n
    flags = capi.MMQ_GIBBS_TRANSPOSED if args.transposed else capi.MMQ_GIBBS_DEFAULT
    small = (4 * h.nnz + 8 * h.m) <= 200e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None

    stream = torch.cuda.Stream(device=dev)
    cid_base = rank * args.fragments
    class_id = None
    if args.layout == "collapsed":   # device order by cost, Philox counters stay canonical (as the host program does)
        from mmseq_b200 import hostlib
        rp_s, col_s, k_s, class_id = hostlib.sort_classes_by_cost(h)
        class_id = class_id + cid_base
        h.row_ptr, h.col, h.k = rp_s, col_s, k_s
    H = capi.Handle(h.row_ptr, h.col, h.k, length, weight=h.w, class_id_base=cid_base, device=local, class_id=class_id)
    H.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [capi.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        H.comm_init(uid[0], rank, world)
        if not args.nccl_only:   # count exchange fused into the Gamma kernel over NVLink peer memory
            hs = [None] * world
            dist.all_gather_object(hs, H.p2p_export())
            H.p2p_attach(hs, rank, world)
    H.init_mu()
    mu0 = H.get_mu()
    cls = H.cls_stats() if h.k is not None else None

    K, W, S = args.steps, args.warmup, SWEEPS_PER_STEP
    L = K + W + 1
    sweep = 0

    def step(timed_flags):
        nonlocal sweep
        if flush_buf is not None:
            with torch.cuda.stream(stream):
                flush_buf.zero_()
        H.gibbs(SEED, sweep, S, stride=S, trace_len=L, flags=timed_flags)
        sweep += S

    for _ in range(W):
        step(flags)
    H.synchronize()
    launches0 = capi.launch_count()
    sampler = ClockSampler(physical_gpu_index(local))
    barrier()
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        step(flags | capi.MMQ_GIBBS_TIME_KERNELS)
    ev1.record(stream)
    H.synchronize()
    barrier()
    sampler.stop_flag = True
    sampler.join()
    launches = capi.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    alloc_ms, alloc_n, gamma_ms, gamma_n = H.kernel_times()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pr = [None] * world   # average kernel times of every rank: shows which rank the others wait for
        dist.all_gather_object(pr, [round(alloc_ms / max(alloc_n, 1), 5), round(gamma_ms / max(gamma_n, 1), 5), round(ms / K, 4)])
        per_rank = {"alloc_ms": [p[0] for p in pr], "gamma_ms": [p[1] for p in pr], "step_ms": [p[2] for p in pr]}
    ms_max = float(t.item())
    sweeps_per_s = K * S / (ms_max / 1000.0)
    m_total = h.m * world  # every rank holds fragments_per_gpu rows (weak scaling)
    value = sweeps_per_s * m_total

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        rp, col, kk, ww, ll, mu_h = pin(h.row_ptr), pin(h.col), pin(h.k), pin(h.w), pin(length), pin(mu0)
        barrier()
        t0 = time.perf_counter()
        H2 = capi.Handle(rp, col, kk, ll, weight=ww, class_id_base=cid_base, device=local, class_id=class_id)   # H2D of the CSR shard
        if world > 1:
            H.comm_move_to(H2)   # the process keeps its NCCL communicator across samples
            if not args.nccl_only:
                hs = [None] * world
                dist.all_gather_object(hs, H2.p2p_export())
                H2.p2p_attach(hs, rank, world)
        H2.set_mu(mu_h)                                                                        # H2D
        H2.gibbs(SEED, 0, K * S, stride=S, trace_len=K, flags=flags)
        mu_out = H2.get_mu()                                                                   # D2H
        tr = H2.get_trace()                                                                    # D2H, n x K doubles
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw.item())
        h2d = rp.nbytes + col.nbytes + (kk.nbytes if kk is not None else 0) + (ww.nbytes if ww is not None else 0) + ll.nbytes + mu_h.nbytes
        d2h = mu_out.nbytes + tr.nbytes
        e2e = {"value": K * S / wall * m_total, "unit": "allocations/s", "sweeps_per_s": K * S / wall,
               "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K), "wall_s": wall,
               "what": "mmq_create(H2D shard) + mmq_set_mu + steps*16 sweeps + mmq_get_mu + mmq_get_trace, pinned host buffers"}
        assert np.isfinite(tr).all() and (tr > 0).any()
        H2.close()
    H.close()

    # ---- roofline of the dominant kernel (k_alloc), device time from CUDA events on its stream
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    b_alloc, b_sweep = algorithmic_bytes(h, n, args.weights, args.transposed, S, args.layout, cls)
    alloc_ms_avg = alloc_ms / max(alloc_n, 1)
    achieved = b_alloc / (alloc_ms_avg * 1e-3) / 1e9 if alloc_n else None
    kernel_name = "k_alloc_seg4" if (args.layout == "perfragment" and not args.transposed) else ("k_alloc_cat" if h.k is None and not args.transposed else "k_alloc")
    if cls and cls["in_use"] and not args.transposed:
        kernel_name = "k_alloc_cls"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.fragments == N_C2 and args.transcripts == T_C2 and not args.weights:
        tj = json.load(open(tpath)).get(kernel_name)
        if tj and tj.get("workload") == f"C2-{args.layout}" and not args.haplo:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]   # per launch, from one ncu --set full capture
    roofline = {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(b_alloc), "avg_launch_ms": alloc_ms_avg, "launches_timed": int(alloc_n),
                "share_of_step": alloc_ms / ms if ms > 0 else None,
                "gamma_avg_launch_ms": gamma_ms / max(gamma_n, 1),
                "sweep_bytes": int(b_sweep), "sweep_gbs": b_sweep * sweeps_per_s / 1e9}
    if per_rank:
        roofline["per_rank"] = per_rank
    if cls:
        roofline["class_plan"] = cls

    line = {
        "metric": "gibbs_hit_class_allocations_per_s", "value": value, "unit": "allocations/s",
        "sweeps_per_s": sweeps_per_s, "fragments_per_s": sweeps_per_s * args.fragments * world,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, h, world), "clocks": sampler.result(),
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "prep": prep,
    }

    # ---- the same sample in the reference's own representation (distinct classes + counts k, what the
    # host program feeds the GPU): class-plan kernel, reported beside the headline (N = 1 only)
    if rank == 0 and world == 1 and args.layout == "perfragment" and not args.no_collapsed and not args.weights and not args.haplo:
        try:
            line["collapsed_layout"] = collapsed_line(args, s, dev, stream, length_full=None)
        except Exception as e:   # an extra: never at the expense of the headline line
            line["collapsed_layout"] = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on the host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        P = orc.Problem(h.row_ptr, h.col, h.k, length)
        threads = orc.max_threads()
        mu_c, _, _ = P.gibbs_gsl(mu0, SEED, 1, threads=threads)
        _, _, sec = P.gibbs_gsl(mu_c, SEED, args.cpu_sweeps, threads=threads)
        cpu_sps = args.cpu_sweeps / sec
        line["cpu_baseline"] = {"value": cpu_sps * h.m, "unit": "allocations/s", "sweeps_per_s": cpu_sps, "cores": threads,
                                "kind": "port", "sample": f"{args.cpu_sweeps} full sweeps of the same shard after 1 warm-up sweep"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
