#!/usr/bin/env python
"""bench.py — stylized target views/s at 256x256 with ~130k Gaussians (BASELINE.json metric), B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE cfg2, "re10k_2v"): one scene of 2x256x256 pixel-aligned Gaussians (P = 131 072, SURVEY §8d
synthetic recipe), 1 target view rasterized per step: preprocess -> tile binning -> per-tile depth radix sort ->
alpha blend.  A step is one pass of that path over one scene; the encoder is not part of the step (round-1 scope:
SURVEY §8 rows a11-a14).  Steps rotate over several resident scenes so that consecutive steps do not reuse L2.

  value      views/s with inputs resident in HBM (CUDA-graph replay of the kernel chain), max-over-ranks time
  e2e        same metric through the serving API `RenderSession.run()` with pinned HOST buffers: H2D of the Gaussians and
             cameras and D2H of the image inside the timed region (`e2e_render_cuda`: the same request through the
             reference-signature eager `render_cuda` call, for comparison)
  cfg2_full / cfg3 / cfg4 / cfg5 / encoder_comparator / pose_align: the other BASELINE.json configurations and
             comparators (bench_legs.py), reported as supplementary keys of the same JSON line
  roofline   blend kernel: algorithmic bytes (40 R + 20 HW + 8 T) / CUDA-event time of that kernel, vs measured HBM peak
  cpu_baseline / --impl reference: the CPU oracle (oracle/raster_oracle.c, a port: the upstream rasterizer is
             un-vendored and has no CPU path) on all host threads.
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
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "stylized target views/sec at 256x256, ~130k Gaussians"
WORKLOAD = "cfg2 re10k_2v raster: 1 scene, v=2 (P=131072 Gaussians), 1 target view 256x256, sh_degree 0, scale-invariant"
HW = 256
V_CTX = 2
V_TGT = 1


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(seed):
    from styl3r_b200 import synthetic as syn
    return syn.make_scene(seed=seed, v=V_CTX, V=V_TGT, hw=HW)


# ----------------------------------------------------------------------------- reference / CPU arm


def oracle_step(scene):
    from oracle import raster_oracle as ro
    outs = []
    for v in range(scene["extrinsics"].shape[0]):
        cam = ro.camera_setup(scene["extrinsics"][v], scene["intrinsics"][v], scene["near"][v], scene["far"][v], True)
        m, c = ro.scale_gaussians(scene["means"], scene["covariances"], cam["scale"])
        shs = np.ascontiguousarray(scene["harmonics"].transpose(0, 2, 1))
        outs.append(ro.forward(m, ro.cov3x3_to_6(c), scene["opacities"], cam["view16"], cam["proj16"], cam["campos"],
                               HW, HW, cam["tanx"], cam["tany"], np.zeros(3, np.float32), shs=shs, deg=0))
    return outs


def cpu_baseline(n_steps=3, warmup=1):
    from oracle import raster_oracle as ro
    scene = make_scene(1234)
    for _ in range(warmup):
        oracle_step(scene)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        oracle_step(scene)
    dt = time.perf_counter() - t0
    return {"value": V_TGT * n_steps / dt, "unit": "views/s", "cores": ro.num_threads(), "kind": "port",
            "sample": f"{n_steps} steps of the same workload (oracle/raster_oracle.c, OpenMP over Gaussians and tiles)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import raster_oracle as ro
    ro.set_num_threads(len(os.sched_getaffinity(0)))  # every host core of the box, also under torchrun (OMP_NUM_THREADS=1)
    scenes = [make_scene(1234 + i) for i in range(2)]
    for i in range(args.warmup):
        oracle_step(scenes[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        oracle_step(scenes[i % 2])
    dt = time.perf_counter() - t0
    val = V_TGT * args.steps / dt
    cb = {"value": val, "unit": "views/s", "cores": ro.num_threads(), "kind": "port",
          "sample": f"{args.steps} full steps (1 view each) on {ro.num_threads()} host threads"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "views/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the un-vendored upstream rasterizer (no CPU path "
                   "exists in the reference); parity unpinned"},
        "cpu_baseline": cb, "e2e": {"value": val, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- our arm


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: pin this rank's threads (and with them the first-touch placement of its pinned host buffers) to
    the NUMA node its GPU hangs off, so that the e2e copies do not cross the socket interconnect.  Returns what was
    done (for the JSON line) or None when the topology cannot be read."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"pci": bdf, "node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"pci": bdf, "node": node, "bound": False}
        os.sched_setaffinity(0, allowed)
        return {"pci": bdf, "node": node, "bound": True, "cpus": len(allowed)}
    except Exception as e:  # not fatal: the run just is not pinned
        return {"bound": False, "error": repr(e)[:120]}


def run_ours(args):
    import torch

    from styl3r_b200 import rasterizer as rz
    from styl3r_b200.decoder import cuda_splatting as cs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    numa = None
    if world > 1:
        numa = bind_to_gpu_numa_node(local)  # before any pinned allocation (single-rank runs keep every core for the CPU legs)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local])

    n_slots = args.slots
    K, Wm = args.steps, max(args.warmup, 3)
    # ---- resident inputs: n_slots different scenes per rank (weak scaling: every rank renders its own scenes)
    slots = []
    for s in range(n_slots):
        sc = make_scene(1234 + 1000 * rank + s)
        t = lambda a: torch.as_tensor(a, device=dev)
        g = dict(means=t(sc["means"])[None], cov=t(sc["covariances"])[None], sh=t(sc["harmonics"])[None],
                 opac=t(sc["opacities"])[None], extr=t(sc["extrinsics"]), intr=t(sc["intrinsics"]),
                 near=t(sc["near"]), far=t(sc["far"]))
        # camera set-up exactly as render_cuda does it (torch, on device, once per slot)
        scale = 1 / g["near"]
        extr = g["extr"].clone()
        extr[:, :3, 3] = extr[:, :3, 3] * scale[:, None]
        fov = cs.get_fov(g["intr"])
        proj_t = cs.get_projection_matrix(g["near"] * scale, g["far"] * scale, fov[:, 0], fov[:, 1]).transpose(1, 2).contiguous()
        view_t = extr.inverse().transpose(1, 2).contiguous()
        full = (view_t @ proj_t).contiguous()
        tensors = (g["means"], g["cov"], g["opac"], g["sh"].reshape(1, -1, 1, 3), None, view_t, full, proj_t,
                   extr[:, :3, 3].contiguous(), (0.5 * fov).tan().contiguous(), scale.contiguous(),
                   torch.zeros(V_TGT, 3, device=dev), torch.zeros(V_TGT, dtype=torch.int32, device=dev))
        P = g["means"].shape[1]
        probe = rz.RasterPlan(tensors, 1, P, V_TGT, HW, HW, 1, 0, 9, 4 * P * V_TGT)
        probe.launch()
        st = probe.ctx.status()
        assert not st["overflow"]
        plan = rz.RasterPlan(tensors, 1, P, V_TGT, HW, HW, 1, 0, 9, int(st["num_instances"] * 1.1) + 1024)
        plan.launch()
        slots.append(dict(g=g, plan=plan, R=st["num_instances"], max_tile=st["max_tile_count"], sc=sc))
    torch.cuda.synchronize()
    resident_mb = sum(s["plan"].state.numel() + 64 * P for s in slots) / 1e6
    # ---- CUDA graphs: one per slot
    graphs = []
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for s in slots:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side):
                s["plan"].launch()
            graphs.append(gph)
    torch.cuda.synchronize()
    n_streams = max(1, min(args.streams, n_slots))
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    main = torch.cuda.current_stream()

    def run_steps(count):
        """`count` steps; step i renders scene slot i % n_slots on stream (i % n_slots) % n_streams, so independent
        scenes overlap on the GPU while a slot's buffers are never used by two streams at once."""
        fork = torch.cuda.Event()
        fork.record(main)
        for st in streams:
            st.wait_event(fork)
        for i in range(count):
            sl = i % n_slots
            with torch.cuda.stream(streams[sl % n_streams]):
                graphs[sl].replay()
        for st in streams:
            j = torch.cuda.Event()
            j.record(st)
            main.wait_event(j)

    run_steps(Wm)
    torch.cuda.synchronize()

    # ---- timed region A: K graph replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        torch.cuda.synchronize()
        e0.record()
        run_steps(K)
        e1.record()
        torch.cuda.synchronize()
        barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * V_TGT * K / (ms / 1e3)
    launches = 5 * K  # preprocess, bin_scan, bin_emit, tile_sort, blend per step

    # ---- region B: per-stage kernel time.  Each stage of each resident scene is captured as a CUDA graph of REP back-to-
    # back launches on one stream and timed with CUDA events around the replay (on that stream): the average launch
    # duration without the host launch gaps that events around single ~10-50 us launches would include (the ncu launch
    # list under profiles/ gives the same shares).  Scenes rotate, so every replay starts from a cold L2 for its scene.
    stages = [("preprocess", rz.STAGE_PREPROCESS), ("bin", rz.STAGE_BIN), ("sort", rz.STAGE_SORT),
              ("blend", rz.STAGE_BLEND)]
    REP = 10
    stage_ms = {}
    n_time = min(n_slots, 8)
    for name, m in stages:
        gs = []
        with torch.cuda.stream(side):
            for s_ in slots[:n_time]:
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side):
                    for _ in range(REP):
                        s_["plan"].launch(m)
                gs.append(gph)
        torch.cuda.synchronize()
        for gph in gs:
            gph.replay()
        torch.cuda.synchronize()
        tot = 0.0
        rounds = max(1, min(20, K // (REP * n_time)))
        for _ in range(rounds):
            for gph in gs:
                e0.record()
                gph.replay()
                e1.record()
                e1.synchronize()
                tot += e0.elapsed_time(e1)
        stage_ms[name] = tot / (rounds * n_time * REP)
        del gs
    R_mean = sum(slots[i % n_slots]["R"] for i in range(K)) / K
    T_tiles = (HW // 16) ** 2
    R_timed = sum(s_["R"] for s_ in slots[:n_time]) / n_time   # the scenes whose blend launches were timed above
    blend_bytes = 40.0 * R_timed + 20.0 * HW * HW * V_TGT + 8.0 * T_tiles * V_TGT
    peak, peak_src = peaks()
    achieved = blend_bytes / (stage_ms["blend"] * 1e-3) / 1e9
    traffic = None
    tj = ROOT / "profiles" / "blend_traffic.json"
    if tj.exists():
        try:
            traffic = json.loads(tj.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- supplementary: the blend kernel at the `value` operating point (blend-only graphs of the resident scenes on the
    # same concurrent streams): a single 256-CTA launch is bound by the serial depth chain of its heaviest tile, not by
    # the machine, so the per-launch figure above understates what the kernel sustains when launches overlap
    blend_graphs = []
    with torch.cuda.stream(side):
        for s_ in slots:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side):
                s_["plan"].launch(rz.STAGE_BLEND)
            blend_graphs.append(gph)
    torch.cuda.synchronize()

    def blend_only(count):
        fork = torch.cuda.Event()
        fork.record(main)
        for st in streams:
            st.wait_event(fork)
        for i in range(count):
            with torch.cuda.stream(streams[(i % n_slots) % n_streams]):
                blend_graphs[i % n_slots].replay()
        for st in streams:
            j = torch.cuda.Event()
            j.record(st)
            main.wait_event(j)

    blend_only(Wm)
    torch.cuda.synchronize()
    e0.record()
    blend_only(K)
    e1.record()
    torch.cuda.synchronize()
    blend_pipe_ms = e0.elapsed_time(e1) / K
    roof_pipe = {"kernel": "s3r_blend_blocks_fwd_kernel", "what": f"blend-only CUDA graphs of the resident scenes on {n_streams} concurrent streams",
                 "kernel_ms_equivalent": blend_pipe_ms, "achieved": blend_bytes / (blend_pipe_ms * 1e-3) / 1e9, "unit": "GB/s",
                 "frac": blend_bytes / (blend_pipe_ms * 1e-3) / 1e9 / peak}

    # ---- e2e through the public API with pinned host buffers
    Ke = min(K, 1000)
    host = []
    for s in slots:
        sc = s["sc"]
        pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
        triu = sc["covariances"][:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]]  # packed cov3D_precomp layout (cuda_splatting.py:118)
        host.append(dict(means=pin(sc["means"][None]), cov=pin(sc["covariances"][None]), cov6=pin(triu[None]), sh=pin(sc["harmonics"][None]),
                         opac=pin(sc["opacities"][None]), extr=pin(sc["extrinsics"]), intr=pin(sc["intrinsics"]),
                         near=pin(sc["near"]), far=pin(sc["far"]), bg=torch.zeros(V_TGT, 3).pin_memory()))
    h2d_bytes = sum(v.numel() * v.element_size() for k_, v in host[0].items() if k_ != "cov")  # the session uploads cov6
    d2h_bytes = V_TGT * 3 * HW * HW * 4

    # public serving API: one RenderSession per resident request buffer = CUDA graph of
    #   pinned host inputs -> H2D -> camera kernel -> raster chain -> D2H -> pinned host image.
    # Independent requests are pipelined over a few streams so the PCIe copies of one overlap the kernels of another.
    from styl3r_b200.decoder import RenderSession
    n_e2e_streams = max(1, min(args.e2e_streams, n_slots))
    e2e_streams = [torch.cuda.Stream() for _ in range(n_e2e_streams)]
    sessions = [RenderSession(dict(extrinsics=h["extr"], intrinsics=h["intr"], near=h["near"], far=h["far"],
                                   background=h["bg"], means=h["means"], covariances=h["cov6"], harmonics=h["sh"],
                                   opacities=h["opac"]), (HW, HW), scale_invariant=True) for h in host]

    def e2e_loop(count):
        fork = torch.cuda.Event()
        fork.record(main)
        for st in e2e_streams:
            st.wait_event(fork)
        for i in range(count):
            sl = i % n_slots
            with torch.cuda.stream(e2e_streams[sl % n_e2e_streams]):
                sessions[sl].run()
        for st in e2e_streams:
            j = torch.cuda.Event()
            j.record(st)
            main.wait_event(j)
        torch.cuda.synchronize()
        for s_ in sessions:
            s_.check()

    # correctness of the serving path vs the eager public call (same inputs)
    with torch.no_grad():
        d0 = {k: v.to(dev) for k, v in host[0].items()}
        ref_color, _ = cs.render_cuda(d0["extr"], d0["intr"], d0["near"], d0["far"], (HW, HW), d0["bg"], d0["means"],
                                      d0["cov"], d0["sh"], d0["opac"], scale_invariant=True,
                                      view_set=torch.zeros(V_TGT, dtype=torch.int32, device=dev))
        sessions[0].run()
        torch.cuda.synchronize()
        assert torch.allclose(sessions[0].color_host.to(dev), ref_color, atol=1e-6), "RenderSession != render_cuda"

    with torch.no_grad():
        e2e_loop(2 * n_slots)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        e2e_loop(Ke)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    e2e_ms = max(e0.elapsed_time(e1), wall * 1e3)
    if dist is not None:
        tms = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms.item())
    e2e_val = world * V_TGT * Ke / (e2e_ms / 1e3)

    # ---- what bounds e2e: the same bytes per request moved by plain pinned copies (no kernels) on the same streams -
    # the host<->device ceiling of this box at this rank count (all ranks copy at once, max over ranks)
    def copy_loop(count):  # the sessions' own pinned buffers and device tensors: exactly the copies a request makes
        for i in range(count):
            sl = i % n_slots
            with torch.cuda.stream(e2e_streams[sl % n_e2e_streams]):
                se = sessions[sl]
                se._copy_in()
                se.color_host.copy_(se._plan.color, non_blocking=True)
        torch.cuda.synchronize()

    copy_loop(8)
    barrier()
    t0 = time.perf_counter()
    copy_loop(200)
    copy_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        tms = torch.tensor([copy_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        copy_ms = float(tms.item())
    pcie_ceiling = world * V_TGT * 200 / (copy_ms / 1e3)

    # ---- the same serving path when one upload feeds several target views (the shape of cfg3 / cfg4: 6 views per scene):
    # the Gaussians cross PCIe once per scene, cameras in / images out per view
    e2e_multi = None
    try:
        V6 = 6
        from styl3r_b200 import synthetic as syn
        sess6 = []
        for i6 in range(min(4, n_slots)):
            sc6 = syn.make_scene(seed=4321 + i6, v=V_CTX, V=V6, hw=HW)
            pin6 = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
            tri6 = sc6["covariances"][:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]]
            h6 = dict(extrinsics=pin6(sc6["extrinsics"]), intrinsics=pin6(sc6["intrinsics"]), near=pin6(sc6["near"]),
                      far=pin6(sc6["far"]), background=torch.zeros(V6, 3).pin_memory(), means=pin6(sc6["means"][None]),
                      covariances=pin6(tri6[None]), harmonics=pin6(sc6["harmonics"][None]), opacities=pin6(sc6["opacities"][None]))
            sess6.append(RenderSession(h6, (HW, HW), scale_invariant=True))
        h2d6 = sum(v.numel() * v.element_size() for v in sess6[0].host.values())

        def loop6(count):
            for i in range(count):
                with torch.cuda.stream(e2e_streams[i % n_e2e_streams]):
                    sess6[i % len(sess6)].run()
            torch.cuda.synchronize()
            for s_ in sess6:
                s_.check()

        n6 = max(40, min(Ke // V6, 200))
        with torch.no_grad():
            loop6(2 * len(sess6))
            barrier()
            t0 = time.perf_counter()
            loop6(n6)
            ms6 = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            tms = torch.tensor([ms6], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms6 = float(tms.item())
        e2e_multi = {"value": world * V6 * n6 / (ms6 / 1e3), "unit": "views/s", "views_per_upload": V6,
                     "h2d_bytes_per_view": h2d6 / V6, "d2h_bytes_per_view": 3 * HW * HW * 4, "requests": n6,
                     "what": "RenderSession.run() with 6 target views per uploaded scene (cfg3 / cfg4 shape): the Gaussians cross "
                             "PCIe once per scene; host wall clock, max over ranks"}
        del sess6
    except Exception as e:  # supplementary
        e2e_multi = {"error": repr(e)[:200]}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_baseline()

    # ---- supplementary (not part of `value`): the encoder that produces the Gaussians, cfg2 shapes, random weights
    enc_info = None
    if rank == 0 and world == 1 and not args.no_encoder:
        try:
            from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, GraphedEncoder, get_encoder
            torch.manual_seed(0)
            enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
            enc = enc.to(dev).eval().to_inference(torch.bfloat16)
            gsel = torch.Generator().manual_seed(1234)
            ctx = {"image": (torch.rand(1, V_CTX, 3, HW, HW, generator=gsel) * 2 - 1).to(dev),
                   "intrinsics": torch.tensor([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1.0]]).expand(1, V_CTX, 3, 3).contiguous().to(dev)}
            sty = {"image": (torch.rand(1, 3, HW, HW, generator=gsel) * 2 - 1).to(dev)}
            fast = GraphedEncoder(enc)
            fast(ctx, sty)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fast(ctx, sty)
            e1.record()
            torch.cuda.synchronize()
            enc_ms = e0.elapsed_time(e1) / 10
            cpu_enc = None
            if not args.no_cpu:  # the reference's PyTorch-CPU encoder path, restated (oracle/encoder_oracle.py), same box
                from oracle import encoder_oracle as eo
                ncores = len(os.sched_getaffinity(0))
                sec = eo.time_encoder(ncores, runs=2)
                cpu_enc = {"value": 1.0 / sec, "unit": "scenes/s", "s_per_scene": sec, "cores": ncores, "kind": "port",
                           "sample": "1 warm-up + 2 forwards of the same cfg1 workload (b=1, v=2, 256x256, random weights), "
                                     "median; oracle/encoder_oracle.py pinned on reference goldens"}
            enc_info = {"ms_per_scene": enc_ms, "tflops": 1270.8 / enc_ms, "cpu_baseline": cpu_enc, "config": "b=1, v=2, 256x256 + style image; bf16 ViT "
                        "trunks (tcgen05 GEMM + attention), bf16 NHWC DPT heads on the tcgen05 implicit-GEMM convolution, "
                        "independent branches on concurrent streams, CUDA-graph replay; random weights",
                        "reference_cpu_s_per_scene_survey_probe": 2.72}
            del enc, fast
        except Exception as e:  # supplementary only
            enc_info = {"error": repr(e)[:200]}

    # ---- supplementary: measured GPU comparator (upstream-style stand-in, baseline/upstream_style) on slot 0
    standin = None
    if rank == 0 and world == 1 and not args.no_standin:
        try:
            from baseline import upstream_style as ups
            g0 = slots[0]["g"]
            bg0 = torch.zeros(V_TGT, 3, device=dev)
            fn = lambda: ups.render_cuda_upstream_style(g0["extr"], g0["intr"], g0["near"], g0["far"], (HW, HW), bg0,
                                                        g0["means"], g0["cov"], g0["sh"], g0["opac"])
            with torch.no_grad():
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(30):
                    fn()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 30
            standin = {"views_per_s": V_TGT / dt, "what": "upstream-style stand-in (per-view launch chain, CUB scan + 64-bit "
                       "radix sort, D2H num_rendered, scalar 16x16 render, reference-style Python loop) on the same "
                       "GPU and inputs; parity unpinned; the un-vendored upstream rasterizer cannot be built here",
                       "value_over_standin": value / (V_TGT / dt)}
        except Exception as e:
            standin = {"error": repr(e)[:200]}

    # ---- supplementary legs: the other BASELINE.json configurations (bench_legs.py); every leg is bounded
    import bench_legs as bl
    extra = {}

    def leg(name, fn, *a, **k):
        if args.legs != "all" and name not in args.legs.split(","):
            return None
        try:
            t_leg = time.perf_counter()
            r = fn(*a, **k)
            r["leg_wall_s"] = round(time.perf_counter() - t_leg, 1)
            return r
        except Exception as e:  # supplementary only: never lose the headline line
            import traceback
            return {"error": repr(e)[:300], "trace": traceback.format_exc()[-600:]}

    if args.legs != "none":
        del graphs, blend_graphs
        for s_ in slots[1:]:
            s_["plan"] = None
        torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            extra["e2e_render_cuda"] = leg("e2e_render_cuda", bl.e2e_render_cuda, dev, host)
            extra["pose_align"] = leg("pose_align", bl.pose_align_leg, dev)
            extra["cfg2_full"] = leg("cfg2_full", bl.full_pipeline, dev, 1, 2, 1, 20,
                                     "cfg2 re10k_2v forward: 2x256x256 in, 131072 Gaussians, 1 target view, 1 scene per pass")
            extra["encoder_comparator"] = leg("encoder_comparator", bl.encoder_comparator, dev)
        # cfg3 / cfg4 / cfg5 run on every rank (cfg4 = cfg3's shape per GPU; cfg5 = DDP training step)
        c3 = leg("cfg3", bl.full_pipeline, dev, 4, 4, 6, 5,
                 "cfg3 re10k_dl3dv_4v forward: 4 scenes x 4x256x256 in, 262144 Gaussians per scene, 6 target views each (24 views per pass)")
        if c3 is not None and "error" not in c3 and dist is not None:
            tms = torch.tensor([c3["ms_per_pass"]], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            c3["ms_per_pass_max_over_ranks"] = float(tms.item())
        if c3 is not None:
            extra["cfg3"] = c3
            if "error" not in c3:
                ms3 = c3.get("ms_per_pass_max_over_ranks", c3["ms_per_pass"])
                extra["cfg4"] = {"workload": f"batched stylized inference: {4 * world} scenes x 6 target views, scene-sharded over {world} GPU(s) "
                                             "(4 scenes = 24 views per GPU per pass, encoder included)", "n_gpus": world,
                                 "views_per_s": world * 24 / (ms3 * 1e-3), "ms_per_pass": ms3}
        barrier()
        c5 = leg("cfg5", bl.train_step_leg, dev, args.train_batch, 3, world, "bf16")
        if world == 1:  # comparator on the same box: the reference's own autograd path (fp32 / TF32 torch ops)
            extra["cfg5_torch_autograd"] = leg("cfg5", bl.train_step_leg, dev, args.train_batch, 3, world, "torch")
        if c5 is not None and "error" not in c5 and dist is not None:
            tms = torch.tensor([c5["ms_per_step"]], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            c5["ms_per_step"] = float(tms.item())
        if c5 is not None:
            if "error" not in c5:
                c5["scenes_per_s"] = world * c5["batch_per_gpu"] / (c5["ms_per_step"] * 1e-3)
                c5["n_gpus"] = world
            extra["cfg5"] = c5

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "P": P, "R_mean": R_mean, "max_tile_instances": max(s["max_tile"] for s in slots),
                       "l2": f"inputs larger than L2: steps rotate over {n_slots} resident scenes ({resident_mb:.0f} MB)",
                       "launch": f"CUDA graph replay of the 5-kernel chain, independent scenes on {n_streams} concurrent streams; "
                                 "camera matrices precomputed per scene",
                       "parallelism": f"scene-sharded x{world}, no collective"},
            "clocks": clk.summary(),
            **({"numa": numa} if numa is not None else {}),
            "e2e": {"value": e2e_val, "unit": "views/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": Ke, "api": "styl3r_b200.decoder.RenderSession.run(): pinned host Gaussians (covariances as the packed cov3D_precomp triangle the reference hands to its rasterizer) + cameras -> H2D -> camera kernel -> "
                           f"raster chain -> D2H pinned image (one CUDA graph per request buffer, {n_e2e_streams} streams)",
                    "pcie_ceiling_views_per_s": pcie_ceiling,
                    "pcie_ceiling_what": "the same H2D + D2H bytes per request as plain pinned copies on the same streams, no kernels, all ranks at "
                                         "once (max over ranks): e2e / ceiling = " + f"{e2e_val / pcie_ceiling:.2f}",
                    "h2d_gbs": e2e_val * h2d_bytes / 1e9},
            "e2e_multi_view": e2e_multi,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "s3r_blend_blocks_fwd_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": blend_bytes, "kernel_ms": stage_ms["blend"],
                         "timing": "CUDA events around a graph of 10 back-to-back launches of the kernel, per launch",
                         "note": "blend is issue-bound (FP32/MUFU) by arithmetic intensity, and a single 256-CTA launch is bound by "
                                 "the serial depth chain of its heaviest tile (DESIGN.md §5); roofline_pipelined = the same "
                                 "kernel with overlapping launches"},
            "roofline_pipelined": roof_pipe,
            "stage_ms": stage_ms,
            "cpu_baseline": cb,
            "encoder": enc_info,
            "upstream_style_standin": standin,
            **{k: v for k, v in extra.items() if v is not None},
        }))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slots", type=int, default=16)
    ap.add_argument("--e2e-streams", type=int, default=4, help="streams pipelining independent e2e requests")
    ap.add_argument("--streams", type=int, default=8, help="concurrent streams over independent scenes")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-standin", action="store_true", help="skip the upstream-style GPU comparator leg")
    ap.add_argument("--with-encoder", action="store_true", help="(default now) time the encoder as a supplementary key")
    ap.add_argument("--no-encoder", action="store_true", help="skip the supplementary encoder leg")
    ap.add_argument("--legs", default="all", help="supplementary legs (bench_legs.py): all | none | comma list of "
                    "e2e_render_cuda,pose_align,cfg2_full,encoder_comparator,cfg3,cfg5")
    ap.add_argument("--train-batch", type=int, default=10, help="scenes per GPU of the cfg5 training-step leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
