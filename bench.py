#!/usr/bin/env python
"""bench.py - molecules/s of the coarse-grained sampling path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--engine strict|fast|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is ONE pass of the hot path over one batch: a full ``sample`` of BASELINE.json configs[1]
(B=64 molecules per GPU, N=40 nodes, 4-layer EGNN, hidden 256, T=1000 -> 1001 EGNN forwards + 1000 diffusion
updates + the final decode).  Prints ONE JSON line (rank 0).  See the repo prompt / DESIGN.md for the keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, N_NODES, N_LAYERS, T_STEPS, HIDDEN = 64, 40, 4, 1000, 256
METRIC = "molecules/sec (batch x N nodes, T=1000)"
UNIT = "molecules/s"


def workload_config(world, engine=None):
    c = {"workload": f"configs[1]: batch={B_PER_GPU}/GPU, N={N_NODES} (all nodes real), T={T_STEPS}, "
                     f"{N_LAYERS}-layer EGNN, hidden={HIDDEN}, attention+tanh, random-init weights",
         "global_batch": B_PER_GPU * world, "n_nodes": N_NODES, "timesteps": T_STEPS, "n_layers": N_LAYERS,
         "parallelism": f"batch-sharded x{world} (no per-step collectives)",
         "cache": "L2 flushed (512 MiB write) between timed steps"}
    if engine:
        c["engine"] = engine
    return c


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path, all host threads
# ------------------------------------------------------------------------------------------------
def oracle_molecules_per_sec(n_mol, n_forwards, seed=0):
    """Time `n_forwards` EGNN forwards (+ diffusion updates) of `n_mol` C2 molecules on the host cores and
    extrapolate to the 1001 forwards of a T=1000 sample.  Returns (mol/s, cores, description)."""
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread (set before libgomp loads)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from oracle import hd_oracle as O
    from weightgen import fill_state_dict
    cfg = O.make_config(N_LAYERS)
    w = O.flatten_weights(cfg, fill_state_dict(O.egnn_shapes(cfg)))
    rng = np.random.default_rng(seed)
    sizes = np.full(n_mol, N_NODES, np.int32)
    z = rng.standard_normal((n_mol, N_NODES, 11)).astype(np.float32)
    z[..., :3] -= z[..., :3].mean(1, keepdims=True)
    rx = rng.standard_normal((n_mol, N_NODES, 3)).astype(np.float32)
    rh = rng.standard_normal((n_mol, N_NODES, 8)).astype(np.float32)
    sc = O.step_scalars(np.float32(2.0), np.float32(2.1))
    O.dynamics_forward(cfg, w, z[:1], np.array([0.5], np.float32), sizes[:1])  # warm the library / threads
    t0 = time.perf_counter()
    for k in range(n_forwards):
        t = np.full(n_mol, 0.5, np.float32)
        eps = O.dynamics_forward(cfg, w, z, t, sizes)
        z = O.reverse_step(z, eps, rx, rh, sizes, sc)
    dt = time.perf_counter() - t0
    per_mol_forward = dt / (n_mol * n_forwards)
    cores = os.cpu_count() or 1
    desc = (f"{n_forwards} reverse steps (EGNN forward + update) of {n_mol} molecules at N={N_NODES}, L={N_LAYERS} "
            f"on the oracle port ({dt:.1f} s, OpenMP on {cores} threads), extrapolated x{T_STEPS + 1} forwards/molecule")
    return 1.0 / (per_mol_forward * (T_STEPS + 1)), cores, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, desc, cores = [], "", 1
    for i in range(args.warmup + args.steps):
        v, cores, desc = oracle_molecules_per_sec(n_mol=16, n_forwards=4, seed=i)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * B_PER_GPU * args.gpus / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is PyTorch-on-CPU and cannot travel to the GPU box; this arm times oracle/ "
                    "(the pinned C restatement of the same arithmetic) on all host threads; each step is a "
                    "bounded sample extrapolated to the full T=1000 chain"}
    print(json.dumps(line), flush=True)


def run_torch_eager(args):
    """GPU comparator of BASELINE.md 4 (">= 10x the reference single-GPU PyTorch"): oracle/torch_port.py issues the
    reference's per-step ATen operator sequence (dense edge index, gathers, cat, Linear, scatter_add_, host syncs)
    in eager PyTorch on cuda:0.  A bounded number of reverse steps is timed and extrapolated to T=1000."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from oracle import torch_port
    dev = torch.device("cuda", 0)
    B, N = B_PER_GPU, N_NODES
    port = torch_port.build(N_LAYERS, dev)
    nm, em = torch_port.masks([N] * B, N, dev)
    out = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.manual_seed(0)
        z = torch.randn(B, N, 11, device=dev)
        z[..., :3] -= z[..., :3].mean(1, keepdim=True)
        t = torch.full((B, 1), 0.5, device=dev)
        sched = (1.0005, 0.01, 0.02)
        n_steps = 10
        with torch.no_grad():
            for _ in range(3):
                port.reverse_step(z, t, sched, nm, em)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n_steps):
                z2 = port.reverse_step(z, t, sched, nm, em)
            b.record()
            torch.cuda.synchronize()
        ms_step = a.elapsed_time(b) / n_steps
        out["tf32" if tf32 else "fp32"] = {"ms_per_reverse_step": ms_step,
                                           "molecules_per_s": B / (ms_step * 1e-3 * (T_STEPS + 1))}
    torch.backends.cuda.matmul.allow_tf32 = False
    line = {"impl": "torch-eager", "metric": METRIC, "value": out["fp32"]["molecules_per_s"], "unit": UNIT,
            "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": workload_config(1), "detail": out,
            "note": "operator-for-operator PyTorch restatement of the reference step (oracle/torch_port.py, pinned "
                    "against the golden fixtures), eager on cuda:0, 10 reverse steps timed with CUDA events and "
                    "extrapolated x1001; 'tf32' = torch.backends.cuda.matmul.allow_tf32 (the reference's torch-1.9 "
                    "default on Ampere)"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from hierdiff_b200 import DiffusionQM9, native, parallel
    from hierdiff_b200.config import default_model_cfg
    import yaml

    ctx = parallel.init()
    if ctx.world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={ctx.world}: launch with torchrun for N>1")
    dev = ctx.device
    hist = os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else ".",
                        f".bench_hist_{ctx.rank}.yaml")
    with open(hist, "w") as f:
        yaml.safe_dump({N_NODES: 1}, f)          # synthetic size distribution: every molecule has N nodes
    torch.manual_seed(2022)                      # identical random-init weights on every rank, then broadcast
    model = DiffusionQM9(default_model_cfg(n_layers=N_LAYERS, timesteps=T_STEPS, analyze=hist)).to(dev).eval()
    os.remove(hist)
    bcast_bytes = parallel.broadcast_parameters(model, ctx)
    engine = args.engine
    if not native.engine_available(engine):
        raise SystemExit(f"engine {engine} is not available in the native library")
    model.engine = engine
    model.steps_per_graph = args.steps_per_graph
    model.use_cuda_graph = not args.no_graph
    torch.manual_seed(ctx.rank)                  # rank r samples with seed r

    B, N = B_PER_GPU, N_NODES
    sizes_pinned = torch.full((B,), N, dtype=torch.int32).pin_memory()
    if args.sizes == "geom":     # secondary workload of SURVEY.md 8d: GEOM size histogram clipped to N, max forced to N
        import numpy as np
        g = np.load(os.path.join(ROOT, "tests", "golden", "nodes_dist.npz"))
        keys, cnt = g["hist_keys"].astype(np.int64), g["hist_counts"].astype(np.float64)
        draw = np.random.default_rng(ctx.rank).choice(keys, size=B, p=cnt / cnt.sum())
        draw = np.minimum(draw, N)
        draw[0] = N
        sizes_pinned = torch.from_numpy(draw.astype(np.int32)).pin_memory()
    loop = model.sampling_loop(B, N, dev)        # builds the schedule table, captures the graph (untimed)
    L = native.lib()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def chain_device():
        loop.run(sizes_pinned)

    def chain_e2e():
        return model.sample_padded(sizes_pinned, dev)

    # launches of OUR kernels per reverse step (eager step outside any timing)
    loop.run(sizes_pinned)
    torch.cuda.synchronize()
    c0 = L.hd_launch_count()
    loop._step()
    torch.cuda.synchronize()
    per_step = L.hd_launch_count() - c0
    c0 = L.hd_launch_count()
    loop._final()
    torch.cuda.synchronize()
    per_final = L.hd_launch_count() - c0

    for _ in range(max(args.warmup - 1, 0)):
        chain_device()
    torch.cuda.synchronize()

    def timed(fn, k):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        parallel.barrier(ctx)
        torch.cuda.synchronize()
        for a, b in ev:
            flush.fill_(1)                       # evict L2 between timed steps (not timed)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        parallel.barrier(ctx)
        ms = sum(a.elapsed_time(b) for a, b in ev)
        return parallel.max_over_ranks(ms, ctx)

    clocks = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        clocks.start()
    ms_dev = timed(chain_device, args.steps)
    ms_e2e = timed(chain_e2e, args.steps)
    clock_info = clocks.stop() if ctx.rank == 0 else None

    # dominant kernel alone: the fused GCL edge kernel of block 0, sub-layer 0 (CUDA events on its stream)
    roof = None
    if ctx.rank == 0:
        roof = edge_kernel_roofline(model, loop, engine, B, N)

    if ctx.rank == 0:
        mols = B * ctx.world * args.steps
        value = mols / (ms_dev / 1e3)
        e2e = mols / (ms_e2e / 1e3)
        cpu = None
        if args.gpus == 1 and not args.no_cpu_baseline:
            v, cores, desc = oracle_molecules_per_sec(n_mol=16, n_forwards=12)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"strict": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate)",
                          "fast": "bf16 operands, fp32 accumulate", "fp32": "f32"}[engine],
                "data": "synthetic", "config": dict(workload_config(ctx.world, engine), sizes=args.sizes,
                                                    mean_nodes=float(sizes_pinned.float().mean())),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(sizes_pinned.numel() * 4),
                        "d2h_bytes_per_step": int(B * N * 11 * 4 + 4), "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int((per_step * T_STEPS + per_final + 1) * args.steps * 2),
                "launches_per_reverse_step": int(per_step), "clocks": clock_info, "roofline": roof,
                "cpu_baseline": cpu, "weight_broadcast_bytes": bcast_bytes,
                "graph": {"steps_per_graph": loop.graph_steps, "enabled": loop.graph is not None}}
        print(json.dumps(line), flush=True)
    parallel.finish(ctx)


def edge_kernel_roofline(model, loop, engine, B, N, iters=20):
    """Time the fused GCL edge kernel alone and express it against the measured tensor / HBM peaks."""
    import torch
    from hierdiff_b200 import native
    L = native.lib()
    egnn = model.dynamics.egnn
    dev = loop.device
    cfg, packed = egnn.hd_config(), egnn.packed_weights()
    ws = egnn.workspace(B, N, dev)
    x = torch.randn(B * N, 3, device=dev)
    h = torch.randn(B * N, HIDDEN, device=dev)
    st = native.stream_ptr()
    eid = native.ENGINES[engine]
    # populate the A|B pre-projection in the workspace
    native.check(L.hd_gcl_forward(cfg, native.ptr(packed), 0, 0, native.ptr(h), native.ptr(x), native.ptr(x),
                                  native.ptr(loop.sizes), B, N, native.ptr(ws), eid, st), "hd_gcl_forward")
    for _ in range(3):
        native.check(L.hd_edge_kernel_only(cfg, native.ptr(packed), 0, 0, native.ptr(x), native.ptr(x),
                                           native.ptr(loop.sizes), B, N, native.ptr(ws), eid, st), "edge_only")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        native.check(L.hd_edge_kernel_only(cfg, native.ptr(packed), 0, 0, native.ptr(x), native.ptr(x),
                                           native.ptr(loop.sizes), B, N, native.ptr(ws), eid, st), "edge_only")
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    n = loop.sizes.double()
    edges = float((n * n).sum())                              # real (i,j) pairs, i == j included (en_dynamics.py:131-136)
    nodes = float(n.sum())
    flops = 2.0 * edges * HIDDEN * HIDDEN                     # dense [E,256]x[256,256] contraction
    hbm_bytes = nodes * (2 * HIDDEN * 4 + HIDDEN * 4 + 24) + 4 * HIDDEN * HIDDEN   # A|B in, agg out, x/x0, W2
    peaks, src = {}, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks, src = json.load(f), "measured"
    except OSError:
        pass
    if engine == "fp32":
        peak, peak_name = 72.0, "nominal fp32 FFMA (148 SM x 128 lanes x 2 x 1.9 GHz)"
    else:
        peak = float(peaks.get("bf16_tflops", 1590.0))
        peak_name = f"bf16 dense burst, {src} (kernel timed alone)"
    traffic = None     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", f"r1_edge_{engine}_metrics.json")) as f:
            m = json.load(f)["launches"][0]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = sum(float(m[k]["value"]) * unit[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except (OSError, KeyError, ValueError, IndexError):
        pass
    # the kernel's second ceiling: every edge-channel needs two SiLU evaluations = 4 MUFU ops (ex2 + rcp, twice) in the
    # strict engine, 2 (tanh, twice) in the fast one, at 16 MUFU lanes / clk / SM
    mufu = (4.0 if engine == "strict" else 2.0 if engine == "fast" else 0.0) * edges * HIDDEN
    props = torch.cuda.get_device_properties(dev)
    sfu_floor_ms = mufu / (16.0 * props.multi_processor_count * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6) * 1e3
    achieved = flops / (ms * 1e-3) / 1e12
    return {"kernel": "fused GCL edge kernel (block 0, gcl_0)", "bound": "tensor", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_name,
            "ms_per_launch": ms, "algorithmic_flops_per_launch": flops,
            "tensor_passes": {"strict": 3, "fast": 1, "fp32": 0}[engine],
            "sfu": {"mufu_ops_per_launch": mufu, "floor_ms": sfu_floor_ms,
                    "frac": (sfu_floor_ms / ms) if mufu else None,
                    "note": "co-bound: SiLU on every edge-channel twice; 16 MUFU lanes/clk/SM at sm_max_mhz"},
            "hbm_GBps_algorithmic": hbm_bytes / (ms * 1e-3) / 1e9,
            "hbm_frac_of_measured": hbm_bytes / (ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
            "note": "SURVEY.md 8d: the fused edge kernel is tensor-pipe bound (AI ~ 32*n FLOP/B); the HBM figure is "
                    "reported beside it because BASELINE.json's metric names it"}


def main():
    global T_STEPS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--engine", default=os.environ.get("HD_BENCH_ENGINE", "strict"), choices=["strict", "fast", "fp32"])
    ap.add_argument("--steps-per-graph", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the loop eagerly (profiling under ncu)")
    ap.add_argument("--sizes", default="full", choices=["full", "geom"],
                    help="full: every molecule has N nodes (headline); geom: sizes from the GEOM histogram, padded to N")
    ap.add_argument("--timesteps", type=int, default=T_STEPS,
                    help="profiling only: a shorter chain (the JSON line then names the shortened workload)")
    args = ap.parse_args()
    T_STEPS = args.timesteps
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-eager":
        run_torch_eager(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
