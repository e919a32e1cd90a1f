#!/usr/bin/env python3
"""bench.py -- decompressed GB/s of the B200 zstd decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload text|literal|single|mixed]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm: oracle port on the host cores

A "step" = one pass of the four-stage decode over one batch of synthetic frames.  At N=1 the
default workload is BASELINE.json configs[1]: 65 536 independent 64 KiB synthetic-text frames
(zstd level 3) decoded in one launch sequence.  With N GPUs every rank decodes its own 65 536
frames (weak scaling; frames shard with a host-side split, no data-path collective).

`value`   : decompressed bytes of all ranks / max-over-ranks device time, inputs + descriptor
            tables resident in HBM, output left in HBM (CUDA events on the launching stream).
`e2e`     : the same through szb_decode_batch with pinned HOST buffers, H2D of the compressed
            bytes and D2H of the decompressed bytes inside the timed region.
`roofline`: algorithmic bytes (C + D + M) / device time vs the measured
            HBM copy bandwidth in MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decompressed_GBps"
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU: run on the CPUs of the GPU's NUMA node, so that the pinned host buffers (first touch) sit in
    the memory the GPU's PCIe root reaches without crossing the socket interconnect.  Best effort; returns the node."""
    try:
        import torch

        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read())
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= set(os.sched_getaffinity(0))
        if node >= 0 and ids:
            os.sched_setaffinity(0, ids)
            return node
    except Exception as e:  # noqa: BLE001
        log(f"[local rank {local_rank}] NUMA binding skipped: {e}")
    return None


def build_corpus(args, rank: int):
    from tools import corpus as cg

    t0 = time.time()
    if args.workload == "text":
        c = cg.config2_text_frames(args.frames, args.frame_size, base_seed=cg.BASE_SEED + rank * args.frames)
    elif args.workload == "literal":
        n = max(1, (args.frames * args.frame_size) // (1 << 20))
        c = cg.config4_literal_heavy(n, 1 << 20, base_seed=cg.BASE_SEED + 10_000_000 + rank * n)
    elif args.workload == "single":
        c = cg.config3_single_frame(args.frames * args.frame_size, 23, seed=cg.BASE_SEED + 20_000_000 + rank)
    elif args.scaling == "strong":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        c = cg.config5_mixed(args.total_bytes, seed=cg.BASE_SEED + 30_000_000, shard=(rank, world))
    else:
        c = cg.config5_mixed(args.frames * args.frame_size, seed=cg.BASE_SEED + 30_000_000 + rank)
    log(f"[rank {rank}] corpus '{c.name}': {c.nframes} frames, C={c.compressed_bytes} D={c.decompressed_bytes} in {time.time() - t0:.1f}s")
    return c


def cpu_reference_arm(args, c, rank, world):
    """--impl reference: the reference decoder's CPU path.  The Go reference cannot be built (no Go
    toolchain in this image), so this times the oracle -- the C restatement of sparkzstd -- with all
    host threads, one decoder per thread over independent frames."""
    from oracle import pyszo
    from tools import corpus as cg

    cores = cg.host_threads()
    # bounded sample of the same workload: sized from a quick probe so one step takes a few seconds
    probe = min(c.nframes, 2 * cores)
    t0 = time.perf_counter()
    pyszo.decode_batch_mt(c.src, c.frame_off[:probe], c.frame_len[:probe], cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = float(c.raw_size[:probe].sum()) / dt
    budget_s = 20.0 / max(1, args.steps + args.warmup)
    nsample = int(min(c.nframes, max(probe, rate * budget_s / max(1.0, float(c.raw_size.mean())))))
    sample_bytes = float(c.raw_size[:nsample].sum())
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, status = pyszo.decode_batch_mt(c.src, c.frame_off[:nsample], c.frame_len[:nsample], cores)
        dt = time.perf_counter() - t0
        assert not status.any()
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample_bytes / (ms * 1e-3) / 1e9
    sample = f"first {nsample} frames of the workload ({sample_bytes / 1e6:.0f} MB decompressed) per step"
    # two more CPU lines SURVEY.md section 8d asks for, each on a couple of seconds of work: the same port on ONE thread, and
    # libzstd 1.5.5 (dlopen) on all threads -- an industrial decoder, NOT the reference, for scale only
    extra = {}
    n1 = int(max(1, min(nsample, nsample * 2.0 / max(1e-3, ms * 1e-3) / max(1, cores))))
    t0 = time.perf_counter()
    _, st1 = pyszo.decode_batch_mt(c.src, c.frame_off[:n1], c.frame_len[:n1], 1)
    dt = time.perf_counter() - t0
    if not st1.any():
        extra["port_1_thread"] = {"value": float(c.raw_size[:n1].sum()) / dt / 1e9, "unit": UNIT, "frames": n1}
    if cg.zstd_available():
        cg.zstd_decode_batch_mt(c, min(nsample, 4 * cores), cores)  # warm
        t0 = time.perf_counter()
        failed = cg.zstd_decode_batch_mt(c, nsample, cores)
        dt = time.perf_counter() - t0
        if failed == 0:
            extra["libzstd_1_5_5"] = {"value": sample_bytes / dt / 1e9, "unit": UNIT, "cores": cores, "frames": nsample,
                                      "note": "not the reference: libzstd via dlopen, for scale"}
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": c.meta.get("workload", c.name), "note": "CPU arm: oracle (C restatement of sparkzstd), one decoder per host thread; Go toolchain absent"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **extra},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def ncu_traffic(workload: str, nframes: int, stage_name: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels of one stage, from the newest
    profiles/*_ncu_traffic.json -- written by scripts/ncu_traffic.py from an `ncu --set full` capture of this command and stamped
    with the commit it was taken on.  None when no capture of this workload and size exists."""
    import glob

    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json"))):
        try:
            with open(f) as fh:
                d = json.load(fh)
        except (OSError, ValueError):
            continue
        if d.get("workload") == workload and int(d.get("frames", -1)) == int(nframes):
            best = (f, d)
    if best is None:
        return None, None
    f, d = best
    names = d.get("stages", {}).get(stage_name, [])
    tot = sum(d["kernels"][k]["dram_read"] + d["kernels"][k]["dram_write"] for k in names if k in d["kernels"])
    if not tot:
        return None, None
    return float(tot), (f"{os.path.relpath(f, ROOT)} (commit {d.get('commit', '?')}): dram__bytes_read.sum + dram__bytes_write.sum "
                        f"per launch of {', '.join(k for k in names if k in d['kernels'])}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="text", choices=["text", "literal", "single", "mixed"])
    ap.add_argument("--frames", type=int, default=65536, help="frames per GPU (text) / size multiplier for other workloads")
    ap.add_argument("--frame-size", type=int, default=65536)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong (mixed workload only): ONE corpus of --total-bytes, frame-sharded over the ranks by szb_shard_frames "
                         "(configs[4] as BASELINE.json words it); weak: every rank decodes its own corpus")
    ap.add_argument("--total-bytes", type=int, default=16 << 30, help="decompressed size of the whole corpus for --scaling strong")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pageable", action="store_true", help="also time the end-to-end call with ordinary (pageable) host buffers: "
                    "the library stages them through its pinned rings (reported as e2e.pageable)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scaling == "strong" and args.workload != "mixed":
        raise SystemExit("--scaling strong is the frame-sharded mixed corpus (configs[4]): use --workload mixed")
    # one process per GPU on a shared host: the library's host-side workers (header walks, staging copies) get cores / ranks
    os.environ.setdefault("SZB_WALK_THREADS", str(max(1, min(8, (os.cpu_count() or 8) // max(world, 1) - 1))))
    # exactly ONE JSON line may reach stdout: libraries (NCCL prints its version on stdout) get stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        if rank != 0:
            return 0
        c = build_corpus(args, 0)
        print(json.dumps(cpu_reference_arm(args, c, rank, world)), file=json_out, flush=True)
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from sparkzstd_b200.decompression import Batch, Context
    from tools import corpus as cg

    c = build_corpus(args, rank)
    ctx = Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # ---- resident inputs: compressed bytes + descriptor tables in HBM ----
    t0 = time.time()
    batch = Batch(ctx, c.src, c.frame_off, c.frame_len)
    walk_s = time.time() - t0
    d_src_t = torch.from_numpy(c.src).to(f"cuda:{local_rank}")
    D = c.decompressed_bytes
    d_dst_t = torch.empty(D + 256, dtype=torch.uint8, device=f"cuda:{local_rank}")
    d_src, d_dst = d_src_t.data_ptr(), d_dst_t.data_ptr()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage_ms = []
    for _ in range(args.warmup):
        batch.run(d_src, d_dst, D + 256)
        st = batch.finish()
        assert not st.any(), f"decode failed: {st[np.nonzero(st)[0][:5]]}"
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            batch.run(d_src, d_dst, D + 256)
        ev1.record(stream)
        ev1.synchronize()
        barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    st = batch.finish()
    assert not st.any()
    stage = ctx.last_timing()  # per-stage device events of the last step
    sizes, _ = batch.read_block_results()

    # algorithmic bytes C + D + M; M = sum of match lengths = regenerated bytes of compressed blocks - their literals
    from sparkzstd_b200.decompression import Walk

    with Walk(c.src, c.frame_off, c.frame_len) as w:
        blocks = w.blocks()
    M = int(sum(int(sizes[i]) - int(b.lit_regen) for i, b in enumerate(blocks) if b.type == 2))
    C_bytes = c.compressed_bytes
    nseq = int(sum(int(b.nseq) for b in blocks))
    lit_huf = int(sum(int(b.lit_regen) for b in blocks if b.type == 2 and b.lit_type >= 2))
    lit_comp = int(sum(int(b.lit_comp) for b in blocks if b.type == 2 and b.lit_type >= 2))

    # which format paths the timed corpus exercises (SURVEY.md section 8d asks for this next to the mixed result)
    cov = {"block_types": [0, 0, 0], "literal_types": [0, 0, 0, 0], "ll_modes": [0, 0, 0, 0], "of_modes": [0, 0, 0, 0], "ml_modes": [0, 0, 0, 0]}
    for b in blocks:
        cov["block_types"][b.type] += 1
        if b.type == 2:
            cov["literal_types"][b.lit_type] += 1
            if b.nseq:
                cov["ll_modes"][b.seq_modes >> 6] += 1
                cov["of_modes"][(b.seq_modes >> 4) & 3] += 1
                cov["ml_modes"][(b.seq_modes >> 2) & 3] += 1
    cov["legend"] = "block types Raw/RLE/Compressed; literal types Raw/RLE/Compressed/Treeless; modes Predefined/RLE/FSE/Repeat"
    if args.workload == "mixed" and args.scaling == "strong" and world > 1:
        # a shard may miss a format path; the corpus as a whole must not: sum the counts over the ranks
        flat = [x for k in ("block_types", "literal_types", "ll_modes", "of_modes", "ml_modes") for x in cov[k]]
        tot_cov = torch.tensor(flat, dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(tot_cov)
        assert bool((tot_cov > 0).all()), f"the sharded mixed corpus misses a format path: {tot_cov.tolist()}"
        cov["all_ranks"] = tot_cov.tolist()
    if args.workload == "mixed" and (args.scaling == "weak" or world == 1):
        assert all(cov["block_types"]) and all(cov["literal_types"]) and all(cov["ll_modes"]) and all(cov["of_modes"]) and \
            all(cov["ml_modes"]), f"the mixed corpus misses a format path: {cov}"

    verified = None
    if not args.no_verify:
        host = d_dst_t[:D].cpu().numpy()
        _, foff, flen = batch.sizes()
        got = cg.hash_frames(host, foff, flen)
        known = c.raw_hash != 0
        verified = bool((flen == c.raw_size).all() and (got[known] == c.raw_hash[known]).all())
        del host
        assert verified, "GPU output does not match the generator's per-frame hashes"

    ms_step = total_ms / args.steps
    t = torch.tensor([ms_step], dtype=torch.float64, device=f"cuda:{local_rank}")
    tot = torch.tensor([float(D), float(C_bytes), float(M)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_step_max = float(t.item())
    D_all, C_all, M_all = [float(x) for x in tot.tolist()]
    shards = None
    if world > 1:  # what every rank had to do and how long it took: the imbalance a strong-scaling split leaves
        mine = torch.tensor([float(D), ms_step], dtype=torch.float64, device=f"cuda:{local_rank}")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        shards = {"decompressed_bytes": [int(e[0].item()) for e in every], "ms_per_step": [float(e[1].item()) for e in every]}
    value = D_all / (ms_step_max * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI batch call, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        h_src = torch.empty(c.src.nbytes, dtype=torch.uint8, pin_memory=True)
        h_src.numpy()[:] = c.src
        h_dst = torch.empty(D + 256, dtype=torch.uint8, pin_memory=True)
        hs, hd = h_src.numpy(), h_dst.numpy()
        e_steps = max(1, min(args.steps, 3))
        ctx.decode_batch_into(hs, c.frame_off, c.frame_len, hd)  # warm-up (allocates the device staging)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(e_steps):
            _, olen, status = ctx.decode_batch_into(hs, c.frame_off, c.frame_len, hd)
        e1.record(stream)
        e1.synchronize()
        barrier()
        assert not status.any() and int(olen.sum()) == D
        e_ms = e0.elapsed_time(e1) / e_steps
        et = torch.tensor([e_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        e2e = {"value": D_all / (float(et.item()) * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(C_bytes),
               "d2h_bytes_per_step": int(D), "ms_per_step": float(et.item()), "includes": "header walk + descriptor upload + H2D + 4 stages + D2H",
               "last_step_ms": ctx.last_timing()}
        if args.e2e_pageable:
            # the same call with malloc'ed buffers (what a Go slice or a Python bytes object is): pinned staging inside the library
            p_src = np.array(c.src, copy=True)
            p_dst = np.empty(D + 256, dtype=np.uint8)
            ctx.decode_batch_into(p_src, c.frame_off, c.frame_len, p_dst)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                _, olen, status = ctx.decode_batch_into(p_src, c.frame_off, c.frame_len, p_dst)
            torch.cuda.synchronize()
            p_ms = (time.perf_counter() - t0) * 1e3 / e_steps
            assert not status.any() and int(olen.sum()) == D
            pt = torch.tensor([p_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            if world > 1:
                dist.all_reduce(pt, op=dist.ReduceOp.MAX)
            e2e["pageable"] = {"value": D_all / (float(pt.item()) * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": float(pt.item()),
                               "note": "host wall clock around the call (it returns when dst is complete); ordinary numpy buffers"}
            del p_src, p_dst
        del h_src, h_dst

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    # Roofline.  ALGORITHMIC bytes only: what a stage must read and write whatever its implementation --
    # intermediate traffic (sequence arrays, literal buffers, segment lists) lowers the achievable fraction, it does not count.
    #   stage 2 (Huffman literals): compressed literal bytes read + regenerated literal bytes written
    #   stage 3 (FSE sequences):    sequences-section bytes read (everything of the compressed blocks that is not literals)
    #   stage 4 (execution):        D + M: decompressed bytes written + match-source bytes read
    # The headline is the PIPELINE fraction (C + D + M) / t_step; `frac` is the same for the stage that takes longest.
    kernels = {
        "k_huffman_literals": {"ms": stage["huffman_literals"], "bytes": lit_comp + lit_huf},
        "k_sequences": {"ms": stage["sequences"], "bytes": C_bytes - lit_comp},
        "k_scan_blocks": {"ms": stage["scan"], "bytes": 16 * len(blocks)},
        "k_execute": {"ms": stage["execute"], "bytes": D + M},
    }
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    dom_ach = kernels[dom]["bytes"] / max(kernels[dom]["ms"], 1e-6) / 1e6
    pipe_ach = (C_bytes + D + M) / (ms_step * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.workload, c.nframes, dom)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": dom_ach, "peak": peak, "unit": "GB/s", "frac": dom_ach / peak, "traffic": traffic,
        "traffic_source": traffic_src, "peak_source": peak_src,
        "kernel_bytes": "algorithmic bytes of that stage per launch (SURVEY.md 8d; DESIGN.md section 4): stage 4 D + M, stage 3 the "
                        "sequences sections, stage 2 compressed + regenerated literals",
        "headline": "pipeline.frac",
        "pipeline": {"algorithmic_bytes": C_bytes + D + M, "C": C_bytes, "D": D, "M": M, "achieved": pipe_ach, "frac": pipe_ach / peak,
                     "frac_of_8TBps_nominal": pipe_ach / 8000.0},
        "stages_ms": {k: v["ms"] for k, v in kernels.items()},
        "stage4_ms": {"k_resolve": stage.get("resolve", 0.0), "rest": stage["execute"] - stage.get("resolve", 0.0)},
        "stage4_path": os.environ.get("SZB_EXEC", "default"),
        "stages_note": "k_huffman_literals and k_sequences run side by side on two streams; each is measured from the start of the step; "
                       "k_execute is all of stage 4: (k_place_zero, k_resolve,) k_frame_verdict, k_execute_bodies, k_execute or k_place and, "
                       "for frames with >= 65 536 sequences, k_execute_pair2 (k_execute_pair from 2 GiB) or the block-parallel kernels k_long_* (execute_long.cuh), "
                       "whichever the host picked for the batch",
    }

    cpu = None
    if not args.no_cpu and world == 1:  # rank 0 at N=1 only
        r = cpu_reference_arm(argparse.Namespace(**{**vars(args), "steps": 2, "warmup": 1}), c, 0, 1)
        cpu = r["cpu_baseline"]

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": c.meta.get("workload", c.name), "frames_per_gpu": c.nframes, "blocks_per_gpu": len(blocks),
                   "sequences_per_gpu": nseq, "compressed_bytes_per_gpu": C_bytes, "decompressed_bytes_per_gpu": D,
                   "parallelism": f"frame-sharded x{world}, no collective",
                   "l2": "inputs (compressed + scratch + output, > 5 GB) are larger than the 126 MB L2; no flush needed",
                   "header_walk_s": walk_s, "format_coverage_per_gpu": cov, "numa_node_rank0": numa,
                   "shards": shards, "shard_plan": c.meta.get("shard")},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "verified": verified,
    }
    print(json.dumps(out), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
