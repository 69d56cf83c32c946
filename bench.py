#!/usr/bin/env python
"""bench.py -- GIGA dense-inference hot path on B200 (BASELINE.json: scenes/s & query-points/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): per GPU a batch of
32 scenes, each a 40^3 TSDF with 2048 grasp query points (quality/rotation/width heads) and 2048
occupancy query points (TSDF head); weak scaling (every rank gets its own 32 scenes), scenes shard
with no data-path collective; the only collective is the NCCL all-gather of the per-scene
(best quality, arg-max index) -- the final grasp-score reduction.

One "step" = one pass of the hot path over one batch: encode (fused conv3d+plane means, U-Net) ->
decode grasp heads -> decode TSDF head -> per-scene arg-max (-> all-gather when N>1).

  value     scenes/s, whole job, inputs resident in HBM, timed with CUDA events on the launch stream
  e2e       same metric through the host entry point (giga_forward_host): pinned HOST buffers,
            H2D + compute + D2H every step
  roofline  dominant kernel (by CUDA-event time; a second, event-instrumented pass of the same K steps) vs the measured bf16 tensor peak
  cpu_baseline  the CPU oracle (port of the reference PyTorch path) on this box's host cores

--impl reference times the reference's own CPU implementation of the path (the oracle port: the
reference is Python/PyTorch and cannot travel to the GPU box) on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scenes_per_sec"
UNIT = "scenes/s"
GRID3 = 40 ** 3
ENC_MAC = 566_476_800                      # per scene (probed)
HEAD_MAC = {"qual": 25_728, "rot": 25_824, "width": 25_728, "tsdf": 25_728}   # per point


def conv_mac_per_plane():
    """MACs per plane image of every encoder kernel (keys = kernel names of the timing report)."""
    m = {}
    conv = {"d0c1": (40, 32, 32), "d0c2": (40, 32, 32), "d1c1": (20, 32, 64), "d1c2": (20, 64, 64), "d2c1": (10, 64, 128),
            "d2c2": (10, 128, 128), "u0c1": (20, 128, 64), "u0c2": (20, 64, 64), "u1c1": (40, 64, 32), "u1c2": (40, 32, 32)}
    for k, (hw, ci, co) in conv.items():
        m["conv3x3:" + k] = hw * hw * ci * co * 9
    m["convT:u0"] = 10 * 10 * 128 * 64 * 4
    m["convT:u1"] = 20 * 20 * 64 * 32 * 4
    m["conv1x1_final"] = 1600 * 32 * 32
    return m


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_cpu_oracle(B, N, steps, warmup, budget_s=25.0):
    """The CPU restatement of the reference path on all host cores; returns (scenes/s, ms/step, steps, cores)."""
    import torch

    from oracle import giga_oracle as O

    ncpu = os.cpu_count() or 1
    sd = O.seeded_state_dict(seed=1)
    x, p, pt = O.seeded_inputs(B, N, seed=0, edge_cases=False)
    # the reference just uses torch's default intra-op pool; on many-core hosts that oversubscribes the
    # small convs, so give the CPU arm its best case: probe a few pool sizes on a 4-scene sample
    cand = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, cores = None, ncpu
    with torch.no_grad():
        for c in cand:
            torch.set_num_threads(c)
            O.forward(sd, x[:4], p[:4], pt[:4])
            t0 = time.perf_counter()
            O.forward(sd, x[:4], p[:4], pt[:4])
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, cores = dt, c
    torch.set_num_threads(cores)
    with torch.no_grad():
        for _ in range(warmup):
            O.forward(sd, x, p, pt)
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            O.forward(sd, x, p, pt)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return B * done / dt, 1e3 * dt / done, done, cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="scenes per GPU")
    ap.add_argument("--points", type=int, default=2048, help="grasp and occupancy query points per scene")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[3] training-step leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    B, N = args.batch, args.points
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"configs[1]: {B} scenes/GPU x (40^3 TSDF, {N} grasp pts x qual/rot/width + {N} occupancy pts x tsdf head)",
              "scenes_per_gpu": B, "grasp_points": N, "occ_points": N, "heads": 4, "parallelism": f"scene-sharded x{world}",
              "l2": "inputs rotate over 16 distinct batches (152 MB > 126 MB L2); ~165 MB of activations rewritten per step"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 10))
        v, ms, done, cores = run_cpu_oracle(B, N, steps, max(1, min(args.warmup, 2)), budget_s=120.0)
        sample = f"{done} steps of the full {B}-scene batch on the host CPU (oracle port of the reference PyTorch path, torch {cores} threads)"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
                          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config, "query_points_per_sec": v * 2 * N,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist

    import giga_b200
    from oracle import giga_oracle as O   # seeded synthetic parameters/inputs + the cpu_baseline leg only

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in giga_b200)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    net = giga_b200.get_network("giga")
    net.load_state_dict(O.seeded_state_dict(seed=1))
    net = net.to(dev)
    POOL = 16
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xs = torch.rand((POOL, B, 40, 40, 40), device=dev, generator=g)
    ps = torch.rand((POOL, B, N, 3), device=dev, generator=g) - 0.5
    pts = torch.rand((POOL, B, N, 3), device=dev, generator=g) - 0.5
    from giga_b200.sharding import SceneBestBuffer
    best = SceneBestBuffer(B, dev)   # [world][2][B]: (best quality, arg-max index) per scene, one all-gather

    def step(i):
        k = i % POOL
        # net(x, p, p_tsdf=...) plus the per-scene arg-max, one C-ABI call (giga_forward)
        (qual, rot, width, occ), _ = net.forward_with_argmax(xs[k], ps[k], pts[k], best.val, best.idx)
        # the final grasp-score reduction: 8 B/scene, one in-place NCCL all-gather on a side stream behind an event (ring of two
        # buffers): the latency-bound exchange of step i overlaps the kernels of step i + 1 (no-op on one GPU)
        best.gather_async()
        return qual, rot, width, occ

    def fence():
        torch.cuda.synchronize()   # all streams of the device, the gather side stream included
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(max(3, args.warmup)):
            step(i)
        fence()
        eng = net._engine()
        sampler = ClockSampler(local_rank)
        sampler.start()
        # ~0.7 s of untimed steps with the sampler running: the clock record then holds >= 5 under-load samples even when the
        # timed K steps themselves last only a few milliseconds
        t_load = time.perf_counter()
        i_load = 0
        while time.perf_counter() - t_load < 0.7:
            for _ in range(20):
                step(i_load)
                i_load += 1
            torch.cuda.synchronize()
        # ---- pass 1 (the reported value): K steps, nothing but the hot path between the two events ----
        l0 = net.gpu_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        fence()
        ms_total = e0.elapsed_time(e1)
        launches = net.gpu_launches - l0
        # ---- pass 2 (roofline breakdown): the same K steps with every kernel bracketed by CUDA events on
        #      the launch stream (giga_ctx_set_timing); shares are relative to this pass's own total ----
        eng.set_timing(True)
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        e2.record()
        for i in range(args.steps):
            step(i)
        e3.record()
        fence()
        ms_total_instr = e2.elapsed_time(e3)
        eng.set_timing(False)
        kern = eng.timing_report()

        # ---- e2e: host buffers in, host buffers out, every step --------------------------------
        hx = [xs[k].cpu().pin_memory() for k in range(2)]
        hp = [ps[k].cpu().pin_memory() for k in range(2)]
        hpt = [pts[k].cpu().pin_memory() for k in range(2)]
        pin = dict(dtype=torch.float32, pin_memory=True)
        hout = (torch.empty((B, N), **pin), torch.empty((B, N, 4), **pin), torch.empty((B, N), **pin), torch.empty((B, N), **pin))
        hout2 = tuple(torch.empty_like(t).pin_memory() for t in hout)
        houts = (hout, hout2)
        for i in range(3):
            net.forward_host(hx[i % 2], hp[i % 2], hpt[i % 2], out=hout)
        for i in range(4):   # warm the pipelined path too (stream / slot creation happens on first use)
            net.forward_host_submit(i % 2, hx[i % 2], hp[i % 2], hpt[i % 2], houts[i % 2])
            net.forward_host_wait(i % 2)
        fence()
        # serving loop through the pipelined host entry point: every step's H2D, kernels and D2H are inside the
        # timed region; step i+1's input copy and step i-1's result copy overlap step i's kernels (two slots)
        e2e_runs = []
        for _rep in range(3):   # K steps each, wall clock around the whole loop; the median run is reported
            fence()
            t0 = time.perf_counter()
            for i in range(args.steps):
                if i >= 2:
                    net.forward_host_wait(i % 2)
                net.forward_host_submit(i % 2, hx[i % 2], hp[i % 2], hpt[i % 2], houts[i % 2])
            for i in range(max(0, args.steps - 2), args.steps):
                net.forward_host_wait(i % 2)
            fence()
            e2e_runs.append(time.perf_counter() - t0)
        e2e_s = sorted(e2e_runs)[1]
        # and the un-pipelined latency of one synchronous host call, for reference
        t1 = time.perf_counter()
        for i in range(5):
            net.forward_host(hx[i % 2], hp[i % 2], hpt[i % 2], out=hout)
        e2e_sync_ms = 1e3 * (time.perf_counter() - t1) / 5
        # ---- configs[2] leg (multi-GPU runs only): 32 scenes / GPU, 4096 grasp + 4096 occupancy points, all four heads ----
        c3 = None
        if world > 1:
            N3 = 4096
            ps3 = torch.rand((4, B, N3, 3), device=dev, generator=g) - 0.5
            pts3 = torch.rand((4, B, N3, 3), device=dev, generator=g) - 0.5

            def step3(i):
                net.forward_with_argmax(xs[i % POOL], ps3[i % 4], pts3[i % 4], best.val, best.idx)
                best.gather_async()

            for i in range(3):
                step3(i)
            fence()
            a3, b3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a3.record()
            for i in range(args.steps):
                step3(i)
            b3.record()
            fence()
            c3 = a3.elapsed_time(b3)
        clocks = sampler.stop()
        # ---- planner: VGNImplicit.__call__ = TSDF in, sorted grasps out, one C-ABI call with host buffers ----
        planner = None
        if world == 1:
            import numpy as np
            from giga_b200.detection_implicit import detect_host, select_params
            from oracle import planner_oracle as PO
            net.load_state_dict(PO.planner_state_dict(O.seeded_state_dict(seed=1)))   # re-centred heads: grasps survive the gates
            tsdfs = np.stack([PO.seeded_volumes(100 + i)[0] for i in range(8)] * (B // 8 if B >= 8 else 1))[:max(B, 1)]
            prm = select_params()
            for nb in (1, len(tsdfs)):
                detect_host(net, tsdfs[:nb], None, prm, K=256)
            l0 = net.gpu_launches
            t0 = time.perf_counter()
            for i in range(20):
                cnt1 = detect_host(net, tsdfs[i % len(tsdfs)][None], None, prm, K=256)[0]
            lat_ms = 1e3 * (time.perf_counter() - t0) / 20
            l1 = net.gpu_launches
            t0 = time.perf_counter()
            for i in range(5):
                cntb = detect_host(net, tsdfs, None, prm, K=256)[0]
            thr = len(tsdfs) * 5 / (time.perf_counter() - t0)
            planner = {"api": "giga_detect_host (H2D 40^3 TSDF, encoder + grasp heads at the 40^3 lattice, smoothing/mask/NMS/sort, D2H of the grasps)",
                       "latency_ms_1_scene": lat_ms, "launches_per_call": (l1 - l0) // 20, "scenes_per_sec_batched": thr, "batch": len(tsdfs),
                       "query_points_per_sec_batched": thr * GRID3, "grasps_found": [int(cnt1[0]), int(cntb.sum())]}
            if not args.no_cpu_baseline:
                with torch.no_grad():
                    q, r, w = net(torch.from_numpy(tsdfs[:1]).to(dev), PO.lattice_positions().to(dev))
                qn, rn, wn = q.cpu().numpy(), r.cpu().numpy(), w.cpu().numpy()
                t0 = time.perf_counter()
                for _ in range(3):
                    PO.detect(tsdfs[:1], qn, rn, wn)
                planner["cpu_postprocess_ms_1_scene"] = 1e3 * (time.perf_counter() - t0) / 3
                planner["cpu_postprocess_note"] = "numpy restatement of scipy.ndimage process/bound/select on the host, post-processing only (no network)"
            net.load_state_dict(O.seeded_state_dict(seed=1))

    # ---- configs[3] leg: a scripts/train_giga.py step (forward + loss + backward + Adam) on this library's kernels only: global batch 64
    #      sharded over the ranks (strong scaling, as the config states), one grasp point + 2048 occupancy points per sample, one flat
    #      NCCL all-reduce of the 581,863-element gradient buffer per step ----
    train_ms = 0.0
    train_info = None
    if not args.no_train and 64 % world == 0:
        import torch.nn.functional as F
        from giga_b200 import training
        TB, TNo = 64 // world, 2048
        tnet = giga_b200.get_network("giga")
        tnet.load_state_dict(O.seeded_state_dict(seed=1))
        tnet = tnet.to(dev)
        opt = training.Adam(tnet.parameters(), lr=2e-4)
        tx = xs[0][:TB] if TB <= B else torch.rand((TB, 40, 40, 40), device=dev, generator=g)
        tpos = torch.rand((TB, 1, 3), device=dev, generator=g) - 0.5
        tocc = torch.rand((TB, TNo, 3), device=dev, generator=g) - 0.5
        ty = ((torch.rand(TB, device=dev, generator=g) > 0.5).float(), F.normalize(torch.randn(TB, 2, 4, device=dev, generator=g), dim=2),
              torch.rand(TB, device=dev, generator=g) * 0.1, (torch.rand(TB, TNo, device=dev, generator=g) > 0.5).float())
        teng = tnet._engine_raw()

        def train_step():
            opt.zero_grad()
            loss, _ = training.loss_fn(training.select(tnet(tx, tpos, p_tsdf=tocc)), ty)
            loss.backward()
            opt.allreduce_gradients()
            opt.step()
            return loss

        for _ in range(3):
            train_step()
        fence()
        tl0 = teng.launches
        ta, tb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ta.record()
        for _ in range(args.steps):
            tloss = train_step()
        tb_.record()
        fence()
        train_ms = ta.elapsed_time(tb_)
        train_info = {"launches_per_step": (teng.launches - tl0) // args.steps, "loss": float(tloss.detach())}
        if world == 1:
            teng.set_timing(True)
            train_step()
            torch.cuda.synchronize()
            rep = teng.timing_report()
            teng.set_timing(False)
            train_info["kernels_us"] = {k: round(1e3 * v[1], 1) for k, v in rep.items()}
        del tnet, opt

    t = torch.tensor([ms_total, e2e_s, c3 or 0.0, train_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, c3, train_ms = t[0].item(), t[1].item(), t[2].item(), t[3].item()
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * args.steps / e2e_s
    h2d = B * GRID3 * 4 + 2 * B * N * 3 * 4
    d2h = B * N * (1 + 4 + 1 + 1) * 4

    if rank == 0:
        peaks = load_peaks()
        macs = conv_mac_per_plane()
        n_img = 3 * B
        table = {}
        for name, (cnt, ms) in kern.items():
            avg_us = 1e3 * ms / cnt
            if name in macs:
                fl = 2.0 * macs[name] * n_img
            elif name in ("conv_in_planes", "conv_in_tc"):
                fl = 2.0 * GRID3 * 27 * 32 * B
            elif name == "decode_points:grasp":
                fl = 2.0 * (HEAD_MAC["qual"] + HEAD_MAC["rot"] + HEAD_MAC["width"]) * B * N
            elif name == "decode_points:grasp+tsdf":
                fl = 2.0 * sum(HEAD_MAC.values()) * B * N
            elif name == "decode_points:tsdf":
                fl = 2.0 * HEAD_MAC["tsdf"] * B * N
            else:
                fl = 0.0
            table[name] = {"us": round(avg_us, 2), "share": round(ms / ms_total_instr, 4), "tflops": round(fl / (avg_us * 1e-6) / 1e12, 3)}
        dom = max(kern, key=lambda k: kern[k][1])
        dom_fl = table[dom]["tflops"]
        sm_clock = clocks.get("sm_mhz") or 1900.0
        fma_peak = 148 * 128 * 2 * sm_clock * 1e6 / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):   # per-launch dram__bytes_read+write of this kernel from the committed ncu --set full capture
            traffic = (json.load(open(tpath)).get(dom) or {}).get("dram_bytes")
        roofline = {"bound": "tensor", "kernel": dom, "achieved": dom_fl, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": round(dom_fl / peaks["bf16_tflops"], 5), "traffic": traffic, "peak_src": peaks["src"],
                    "note": ("dominant kernel by CUDA-event time in the instrumented pass (per-kernel events serialise the programmatic-dependent-launch "
                             "overlap, so the per-kernel times sum to more than ms_per_step); tensor-core kernels run 3xFP16 operand splitting (3 fp16 MMAs "
                             "per fp32-equivalent product, fp32 FLOPs counted once, so the tensor pipe does 3x the reported rate); conv_in_planes is an "
                             "fp32 FMA-pipe kernel: its own roofline is fp32_fma_peak (frac_fp32_fma), `frac` relates it to the tensor peak as the "
                             "contract asks; traffic = ncu dram bytes of that kernel from profiles/traffic.json (same command, committed)"),
                    "alg_bytes_per_launch": {"conv_in_planes": B * (GRID3 * 4 + 3 * 32 * 1600 * 4), "conv_in_tc": B * (GRID3 * 4 + 3 * 32 * 1600 * 4),
                                             "decode_points:grasp+tsdf": B * (3 * 1600 * 32 * 4 + 2 * N * 12 + N * 28)}.get(dom),
                    "fp32_fma_peak": round(fma_peak, 1), "frac_fp32_fma": round(dom_fl / fma_peak, 4),
                    "step_tflops": round(2.0 * (ENC_MAC + N * sum(HEAD_MAC.values())) * B / (ms_per_step * 1e-3) / 1e12, 3),
                    "step_hbm_frac": round((362_496 * B / (ms_per_step * 1e-3)) / 1e9 / peaks["hbm_gbs"], 5)}
        # the dominant kernel flips between conv_in_planes (fp32 FMA pipe) and the decoder (tensor pipe) within measurement noise (~110 us
        # each): report the largest tensor-core kernel beside it so that both pipes' rooflines are on the line whichever one leads
        tc_names = [k for k in kern if k.startswith(("decode_points", "conv3x3", "convT"))]
        if tc_names:
            tk = max(tc_names, key=lambda k: kern[k][1])
            roofline["dominant_tensor_kernel"] = {"kernel": tk, "us": table[tk]["us"], "achieved": table[tk]["tflops"], "unit": "TFLOP/s",
                                                  "frac": round(table[tk]["tflops"] / peaks["bf16_tflops"], 5)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms, done, cores = run_cpu_oracle(B, N, steps=10, warmup=1, budget_s=20.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
                   "sample": f"{done} steps of the full {B}-scene batch (oracle port of the reference PyTorch path, torch {cores} threads)"}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config, "query_points_per_sec": value * 2 * N,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": 1e3 * e2e_s / args.steps, "api": "giga_forward_host_submit/wait (2 slots, pinned host buffers)",
                       "runs_ms_per_step": [round(1e3 * r / args.steps, 4) for r in e2e_runs], "note": "median of 3 runs of K steps",
                       "sync_call_ms": e2e_sync_ms},
               "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
               "kernels_pass": {"ms_per_step_instrumented": ms_total_instr / args.steps, "note": "per-kernel CUDA events (pass 2)"},
               "kernels": table, "planner": planner}
        if world > 1 and c3 > 0:
            out["configs2"] = {"workload": f"configs[2]: {B} scenes/GPU x {world} GPUs, 4096 grasp pts x qual/rot/width + 4096 occupancy pts x tsdf head",
                               "value": world * B * args.steps / (c3 * 1e-3), "unit": UNIT, "ms_per_step": c3 / args.steps,
                               "query_points_per_sec": world * B * args.steps / (c3 * 1e-3) * 2 * 4096}
        if train_info is not None and train_ms > 0:
            # fwd+bwd ~ 3x the forward's FLOPs: 3.72 GFLOP per sample
            tps = 64 * args.steps / (train_ms * 1e-3)
            out["train"] = {"workload": f"configs[3]: train_giga.py step, global batch 64 ({64 // world}/GPU x {world}), 1 grasp pt + 2048 occupancy pts per sample, "
                                        "forward on the tcgen05 kernels (device-packed operands) + fused loss + native backward (fp32 FMA pipe) + flat all-reduce + flat Adam",
                            "value": tps, "unit": "samples/s", "ms_per_step": train_ms / args.steps, "scaling": "strong",
                            "tflops_fp32": round(3.72e9 * tps / 1e12, 2), **train_info}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
