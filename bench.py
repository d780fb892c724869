#!/usr/bin/env python
"""bench.py - molecules/s of the coarse-grained sampling path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch-eager]
                    [--engine strict|fast|fp32] [--workload configs1|geom9|sweep]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is ONE pass of the hot path over one batch: a full ``sample`` (T=1000 -> 1001 EGNN forwards + 1000
diffusion updates + the final decode).  Workloads:

    configs1 (default)  BASELINE.json configs[1]: B=64 molecules per GPU, N=40 nodes (all real), 4-layer EGNN, H=256
    geom9               BASELINE.json configs[4]: 9-layer EGNN, B=64, sizes drawn from the GEOM histogram
    sweep               BASELINE.json configs[2]: B=128, N in {16,24,32,40,56}, 4-layer EGNN (one JSON line per N)

Prints ONE JSON line per workload instance (rank 0).  See the repo prompt / DESIGN.md for the keys.

Reference legs (test/baseline infrastructure, the only places that touch ``oracle/``):
  * ``--impl reference``: the UNMODIFIED reference (``oracle/_ref``, staged by ``oracle/stage_ref.py``) - its own
    ``DiffusionQM9.sample_p_zs_given_zt`` on the host cores (all threads), a bounded number of reverse steps per bench
    step, extrapolated to the 1001 forwards of the chain.  Falls back to the C restatement (``oracle/hd_oracle.c``,
    ``kind: "port"``) only when the staged copy is missing.
  * default arm: ``cpu_baseline`` (same measurement, one bounded sample) and ``gpu_reference`` = the same reference
    class running its own ``sample()`` on cuda:0, whole T=1000 chain, torch fp32 and allow_tf32 - the comparator of
    BASELINE.json's ">= 10x the reference single-GPU PyTorch" target.
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

T_STEPS, HIDDEN = 1000, 256
METRIC = "molecules/sec (batch x N nodes, T=1000)"
UNIT = "molecules/s"
SWEEP_N = [16, 24, 32, 40, 56]
# MUFU operations the edge kernel issues per edge-channel (two SiLU evaluations): strict = 2 x (ex2 + half a rcp: one
# reciprocal serves a pair of elements), fast = 2 x tanh
MUFU_PER_EDGE_CHANNEL = {"strict": 3.0, "fast": 2.0, "fp32": 0.0}


class Workload:
    def __init__(self, name, B, N, L, sizes, label):
        self.name, self.B, self.N, self.L, self.sizes, self.label = name, B, N, L, sizes, label


def geom_sizes(B, seed):
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "nodes_dist.npz"))
    keys, cnt = g["hist_keys"].astype(np.int64), g["hist_counts"].astype(np.float64)
    return np.random.default_rng(seed).choice(keys, size=B, p=cnt / cnt.sum()).astype(np.int32)


def make_workloads(args, rank=0):
    import numpy as np
    if args.workload == "configs1":
        sizes = np.full(64, 40, np.int32)
        label = "configs[1]: batch=64/GPU, N=40 (all nodes real), T=%d, 4-layer EGNN, hidden=256" % T_STEPS
        if args.sizes == "geom":   # secondary workload of SURVEY.md 8d: GEOM sizes clipped to N, max forced to N
            sizes = np.minimum(geom_sizes(64, rank), 40).astype(np.int32)
            sizes[0] = 40
            label = "configs[1] shape with GEOM-histogram sizes clipped to N=40: batch=64/GPU, T=%d, 4-layer EGNN" % T_STEPS
        return [Workload("configs1", 64, 40, 4, sizes, label + ", attention+tanh, random-init weights")]
    if args.workload == "geom9":
        sizes = geom_sizes(64, rank)
        return [Workload("geom9", 64, int(sizes.max()), 9, sizes,
                         "configs[4]: GEOM-drugs config, 9-layer EGNN, hidden=256, T=%d, batch=64/GPU, sizes drawn "
                         "from the GEOM histogram (padded to the batch maximum), random-init weights" % T_STEPS)]
    return [Workload("sweep_n%d" % n, 128, n, 4, np.full(128, n, np.int32),
                     "configs[2]: node-count sweep, batch=128/GPU, N=%d (all nodes real), T=%d, 4-layer EGNN, "
                     "hidden=256, random-init weights" % (n, T_STEPS)) for n in SWEEP_N]


def workload_config(w, world, engine=None):
    c = {"workload": w.label, "global_batch": w.B * world, "n_nodes": w.N, "timesteps": T_STEPS, "n_layers": w.L,
         "mean_nodes": float(w.sizes.mean()),
         "parallelism": f"batch-sharded x{world} (no per-step collectives)",
         "cache": "L2 flushed (512 MiB write) between timed steps"}
    if engine:
        c["engine"] = engine
    return c


# ------------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference (oracle/_ref) or, failing that, the C restatement, on all host threads
# ------------------------------------------------------------------------------------------------
def _all_threads():
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)     # torchrun exports OMP_NUM_THREADS=1
    os.environ["MKL_NUM_THREADS"] = str(n)
    return n


def reference_cpu_molecules_per_sec(w, n_steps, model_cache={}):
    """Time ``n_steps`` reverse steps (``sample_p_zs_given_zt``: EGNN forward + diffusion update, gamma calls, host
    syncs) of the UNMODIFIED reference on the whole batch of workload ``w`` with every host thread, extrapolated to
    the 1001 forwards of a T=1000 sample.  Returns (mol/s, cores, description)."""
    cores = _all_threads()
    import torch
    from oracle import ref_runner as R
    torch.set_num_threads(cores)
    key = (w.L,)
    if key not in model_cache:
        model_cache[key] = R.make_reference(w.L, T_STEPS)
    ref = model_cache[key]
    B, N = w.B, w.N
    dev = torch.device("cpu")
    sizes = [int(v) for v in w.sizes]
    node_mask = torch.zeros(B, N, 1)
    edge_mask = torch.zeros(B, N, N)
    for i, n in enumerate(sizes):
        node_mask[i, :n] = 1
        edge_mask[i, :n, :n] = 1 - torch.eye(n)
    node_mask, edge_mask = node_mask.bool(), edge_mask.bool()
    torch.manual_seed(0)
    with torch.no_grad():
        z = ref.sample_combined_position_feature_noise(B, N, node_mask)
        s0 = T_STEPS - 1

        def step(z, s):
            s_arr = torch.full((B, 1), fill_value=s, device=dev)
            return ref.sample_p_zs_given_zt(s_arr / T_STEPS, (s_arr + 1) / T_STEPS, z, node_mask, edge_mask, None,
                                            mol_shape=N)
        z = step(z, s0)                      # warm the thread pool / allocator
        t0 = time.perf_counter()
        for k in range(n_steps):
            z = step(z, s0 - 1 - k)
        dt = time.perf_counter() - t0
    per_step = dt / n_steps
    desc = (f"{n_steps} reverse steps of the unmodified reference (oracle/_ref: DiffusionQM9.sample_p_zs_given_zt, "
            f"torch {torch.__version__} CPU fp32) on the full batch B={B}, N={N}, L={w.L}: {per_step:.2f} s per step on "
            f"{cores} threads, extrapolated x{T_STEPS + 1} forwards per sample")
    return B / (per_step * (T_STEPS + 1)), cores, desc


def oracle_molecules_per_sec(w, n_mol, n_forwards, seed=0):
    """The C restatement (oracle/hd_oracle.c, OpenMP over molecules) - used only when oracle/_ref is not staged."""
    cores = _all_threads()
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from oracle import hd_oracle as O
    from weightgen import fill_state_dict
    cfg = O.make_config(w.L)
    wts = O.flatten_weights(cfg, fill_state_dict(O.egnn_shapes(cfg)))
    rng = np.random.default_rng(seed)
    sizes = np.ascontiguousarray(w.sizes[:n_mol])
    N = w.N
    z = rng.standard_normal((n_mol, N, 11)).astype(np.float32)
    z[..., :3] -= z[..., :3].mean(1, keepdims=True)
    rx = rng.standard_normal((n_mol, N, 3)).astype(np.float32)
    rh = rng.standard_normal((n_mol, N, 8)).astype(np.float32)
    sc = O.step_scalars(np.float32(2.0), np.float32(2.1))
    O.dynamics_forward(cfg, wts, z[:1], np.array([0.5], np.float32), sizes[:1])
    t0 = time.perf_counter()
    for k in range(n_forwards):
        eps = O.dynamics_forward(cfg, wts, z, np.full(n_mol, 0.5, np.float32), sizes)
        z = O.reverse_step(z, eps, rx, rh, sizes, sc)
    dt = time.perf_counter() - t0
    desc = (f"{n_forwards} reverse steps of {n_mol} molecules at N={N}, L={w.L} on the oracle port "
            f"({dt:.1f} s, OpenMP on {cores} threads), extrapolated x{T_STEPS + 1} forwards/molecule")
    return n_mol * n_forwards / dt / (T_STEPS + 1), cores, desc


def cpu_leg(w, n_steps):
    """(value, cores, kind, sample description) of the CPU baseline for workload ``w``."""
    from oracle import ref_runner as R
    if R.available():
        v, cores, desc = reference_cpu_molecules_per_sec(w, n_steps)
        return v, cores, "reference", desc
    v, cores, desc = oracle_molecules_per_sec(w, n_mol=min(16, w.B), n_forwards=4 * n_steps)
    return v, cores, "port", desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for w in make_workloads(args):
        vals, desc, cores, kind = [], "", 1, "reference"
        for i in range(args.warmup + args.steps):
            v, cores, kind, desc = cpu_leg(w, n_steps=args.ref_steps)
            if i >= args.warmup:
                vals.append(v)
        value = sum(vals) / len(vals)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * w.B / value,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(w, args.gpus),
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
                                 "extrapolated": True},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "the reference's own CPU PyTorch path on this box's host cores (one box, whatever --gpus says); "
                        "each bench step times a bounded number of reverse steps of the full batch and extrapolates "
                        "to the T=1000 chain (a full chain is ~45 min of CPU)"}
        print(json.dumps(line), flush=True)


def gpu_reference(w, dev, tf32, full_chain=True):
    """The UNMODIFIED reference class running its own ``sample()`` on the GPU (torch eager): whole T=1000 chain of
    workload ``w`` timed with CUDA events around the call (it ends with the reference's own per-molecule .cpu())."""
    import torch
    from oracle import ref_runner as R
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        T = T_STEPS if full_chain else 50
        ref = R.make_reference(w.L, T).to(dev)
        ref.nodes_dist = R.FixedNodes(w.sizes)
        warm = R.make_reference(w.L, 3).to(dev)
        warm.load_state_dict(ref.state_dict())
        warm.nodes_dist = R.FixedNodes(w.sizes)
        torch.manual_seed(0)
        warm.sample(w.B, dev)                     # cuBLAS handles, allocator, the reference's cached edge lists
        ref.dynamics._edges_dict = warm.dynamics._edges_dict
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = ref.sample(w.B, dev)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        assert len(res) == w.B
        ms_chain = ms if full_chain else ms * (T_STEPS + 1) / (T + 1)
        return {"molecules_per_s": w.B / (ms_chain * 1e-3), "ms_per_sample_call": ms_chain,
                "ms_per_reverse_step": ms / (T + 1), "timesteps_run": T, "extrapolated": not full_chain}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def run_torch_eager(args):
    """Kept for comparison with round 1: oracle/torch_port.py (a restatement of the reference's operator sequence)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from oracle import torch_port
    dev = torch.device("cuda", 0)
    w = make_workloads(args)[0]
    B, N = w.B, w.N
    port = torch_port.build(w.L, dev)
    nm, em = torch_port.masks([int(v) for v in w.sizes], N, dev)
    out = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.manual_seed(0)
        z = torch.randn(B, N, 11, device=dev) * nm
        z[..., :3] -= (z[..., :3].sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * nm
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
                port.reverse_step(z, t, sched, nm, em)
            b.record()
            torch.cuda.synchronize()
        ms_step = a.elapsed_time(b) / n_steps
        out["tf32" if tf32 else "fp32"] = {"ms_per_reverse_step": ms_step,
                                           "molecules_per_s": B / (ms_step * 1e-3 * (T_STEPS + 1))}
    torch.backends.cuda.matmul.allow_tf32 = False
    line = {"impl": "torch-eager", "metric": METRIC, "value": out["fp32"]["molecules_per_s"], "unit": UNIT,
            "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, 1), "detail": out,
            "note": "operator-for-operator PyTorch restatement of the reference step (oracle/torch_port.py), eager on "
                    "cuda:0, 10 reverse steps timed and extrapolated x1001; superseded by `gpu_reference` of the "
                    "default arm, which runs the unmodified reference itself"}
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


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import yaml
    from hierdiff_b200 import DiffusionQM9, native, parallel
    from hierdiff_b200.config import default_model_cfg

    ctx = parallel.init()
    if ctx.world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={ctx.world}: launch with torchrun for N>1")
    dev = ctx.device
    engine = args.engine
    if not native.engine_available(engine):
        raise SystemExit(f"engine {engine} is not available in the native library")
    L = native.lib()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    models = {}
    for w in make_workloads(args, ctx.rank):
        if w.L not in models:
            hist = os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else ".",
                                f".bench_hist_{ctx.rank}.yaml")
            with open(hist, "w") as f:
                yaml.safe_dump({40: 1}, f)               # sizes are given explicitly below (sample_padded)
            torch.manual_seed(2022)                      # identical random-init weights on every rank, then broadcast
            model = DiffusionQM9(default_model_cfg(n_layers=w.L, timesteps=T_STEPS, analyze=hist)).to(dev).eval()
            os.remove(hist)
            bcast_bytes = parallel.broadcast_parameters(model, ctx)
            model.engine = engine
            model.steps_per_graph = args.steps_per_graph
            model.use_cuda_graph = not args.no_graph
            models[w.L] = (model, bcast_bytes)
        model, bcast_bytes = models[w.L]
        torch.manual_seed(ctx.rank)                  # rank r samples with seed r
        B, N = w.B, w.N
        sizes_pinned = torch.from_numpy(w.sizes.copy()).pin_memory()
        loop = model.sampling_loop(B, N, dev)        # builds the schedule table, captures the graph (untimed)

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

        if ctx.rank == 0:
            roof = edge_kernel_roofline(model, loop, engine, B, N)
            others = [reverse_step_roofline(loop), node_gemm_roofline(model, loop, engine, B, N)]
            mols = B * ctx.world * args.steps
            value = mols / (ms_dev / 1e3)
            e2e = mols / (ms_e2e / 1e3)
            cpu, gpu_ref = None, None
            if args.gpus == 1 and not args.no_cpu_baseline:
                v, cores, kind, desc = cpu_leg(w, n_steps=args.cpu_steps)
                cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc, "extrapolated": True}
            if args.gpus == 1 and not args.no_gpu_reference:
                from oracle import ref_runner as R
                if R.available():
                    fp32 = gpu_reference(w, dev, tf32=False, full_chain=not args.short_gpu_reference)
                    tf32 = gpu_reference(w, dev, tf32=True, full_chain=not args.short_gpu_reference)
                    gpu_ref = {"what": "the unmodified reference (oracle/_ref) DiffusionQM9.sample() on cuda:0, eager "
                                       "torch, same workload, whole chain incl. its per-molecule .cpu()",
                               "fp32": fp32, "tf32": tf32, "unit": UNIT,
                               "ours_e2e_over_reference_fp32": e2e / fp32["molecules_per_s"],
                               "ours_e2e_over_reference_tf32": e2e / tf32["molecules_per_s"]}
                else:
                    gpu_ref = {"unavailable": "oracle/_ref is not staged"}
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None,
                    "dtype": {"strict": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate)",
                              "fast": "bf16 operands, fp32 accumulate", "fp32": "f32"}[engine],
                    "data": "synthetic", "config": dict(workload_config(w, ctx.world, engine), sizes=args.sizes),
                    "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(sizes_pinned.numel() * 4),
                            "d2h_bytes_per_step": int(B * N * 11 * 4 + 4), "ms_per_step": ms_e2e / args.steps},
                    "gpu_launches": int((per_step * T_STEPS + per_final + 1) * args.steps * 2),
                    "launches_per_reverse_step": int(per_step), "clocks": clock_info, "roofline": roof,
                    "roofline_other": others, "cpu_baseline": cpu, "gpu_reference": gpu_ref,
                    "weight_broadcast_bytes": bcast_bytes,
                    "graph": {"steps_per_graph": loop.graph_steps, "enabled": loop.graph is not None}}
            print(json.dumps(line), flush=True)
        model._loops.clear()
    parallel.finish(ctx)


def _time_launches(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def _profile_metrics(name):
    """dram bytes per launch from the committed ncu capture of this round (profiles/r2_*), else round 1's."""
    for rnd in ("r2", "r1"):
        try:
            with open(os.path.join(ROOT, "profiles", f"{rnd}_{name}_metrics.json")) as f:
                m = json.load(f)["launches"][0]
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            return sum(float(m[k]["value"]) * unit[m[k]["unit"]]
                       for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")), f"profiles/{rnd}_{name}_metrics.json"
        except (OSError, KeyError, ValueError, IndexError):
            continue
    return None, None


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
    ms = _time_launches(lambda: native.check(L.hd_edge_kernel_only(
        cfg, native.ptr(packed), 0, 0, native.ptr(x), native.ptr(x), native.ptr(loop.sizes), B, N, native.ptr(ws),
        eid, st), "edge_only"), iters)
    n = loop.sizes.double()
    edges = float((n * n).sum())                              # real (i,j) pairs, i == j included (en_dynamics.py:131-136)
    nodes = float(n.sum())
    flops = 2.0 * edges * HIDDEN * HIDDEN                     # dense [E,256]x[256,256] contraction
    hbm_bytes = nodes * (2 * HIDDEN * 4 + HIDDEN * 4 + 24) + 4 * HIDDEN * HIDDEN   # A|B in, agg out, x/x0, W2
    peaks, src = measured_peaks()
    if engine == "fp32":
        peak, peak_name = 72.0, "nominal fp32 FFMA (148 SM x 128 lanes x 2 x 1.9 GHz)"
    else:
        peak = float(peaks.get("bf16_tflops", 1590.0))
        peak_name = f"bf16 dense burst, {src} (kernel timed alone)"
    traffic, traffic_src = _profile_metrics(f"edge_{engine}")
    # the kernel's second ceiling: SiLU on every edge-channel twice (MUFU_PER_EDGE_CHANNEL); 16 MUFU lanes / clk / SM
    mufu = MUFU_PER_EDGE_CHANNEL[engine] * edges * HIDDEN
    props = torch.cuda.get_device_properties(dev)
    sfu_floor_ms = mufu / (16.0 * props.multi_processor_count * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6) * 1e3
    achieved = flops / (ms * 1e-3) / 1e12
    return {"kernel": "fused GCL edge kernel tc::edge_tc_k (block 0, gcl_0)", "bound": "tensor", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_name, "ms_per_launch": ms, "algorithmic_flops_per_launch": flops,
            "tensor_passes": {"strict": 3, "fast": 1, "fp32": 0}[engine],
            # context for `frac` (which counts the algorithmic one-pass FLOPs): the bf16x3 strict engine issues three
            # tcgen05 passes per product, so its tensor pipe is busy for passes x achieved
            "tensor_issued": ({"tflops": achieved * {"strict": 3, "fast": 1}[engine],
                               "frac_of_peak": achieved * {"strict": 3, "fast": 1}[engine] / peak,
                               "note": "profiles/r2_notes.md: with the operand producers switched off the strict kernel "
                                       "takes 39.8 us (6 tile slots x 4.55 us at the measured bf16 rate + fill/drain), "
                                       "which caps the one-pass fraction near 0.21"}
                              if engine in ("strict", "fast") else None),
            "sfu": {"mufu_ops_per_launch": mufu, "floor_ms": sfu_floor_ms,
                    "frac": (sfu_floor_ms / ms) if mufu else None,
                    "note": "co-bound: SiLU on every edge-channel twice; 16 MUFU lanes/clk/SM at sm_max_mhz"},
            "hbm_GBps_algorithmic": hbm_bytes / (ms * 1e-3) / 1e9,
            "hbm_frac_of_measured": hbm_bytes / (ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
            "note": "SURVEY.md 8d: the fused edge kernel is tensor-pipe bound (AI ~ 32*n FLOP/B); the HBM figure is "
                    "reported beside it because BASELINE.json's metric names it"}


def reverse_step_roofline(loop):
    """The diffusion update (176*N bytes per molecule: z, eps, noise in, z out) against HBM."""
    import torch
    from hierdiff_b200 import native
    L = native.lib()
    B, N, F = loop.B, loop.N, loop.F
    st = native.stream_ptr()
    zs, eps = torch.empty_like(loop.z), torch.randn_like(loop.z)
    sched = loop.table.sched[0].contiguous()
    ms = _time_launches(lambda: native.check(L.hd_reverse_step(
        native.ptr(loop.z), native.ptr(eps), native.ptr(loop.rx), native.ptr(loop.rh), native.ptr(loop.sizes), B,
        N, F, native.ptr(sched), 1, native.ptr(zs), None, st), "hd_reverse_step"))
    peaks, src = measured_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    nbytes = 4.0 * (3 + F) * N * B * 4
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "diffusion update hd::reverse_step_k (timed alone, back to back; in the loop it is part of "
                      "sampler_tail_k)", "bound": "hbm",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": f"copy bandwidth, {src}", "ms_per_launch": ms, "algorithmic_bytes_per_launch": nbytes,
            "note": "176*N bytes per molecule: at B*N = %d nodes the launch moves %.0f KB, so it is launch-latency "
                    "bound, not bandwidth bound" % (B * N, nbytes / 1e3)}


def node_gemm_roofline(model, loop, engine, B, N):
    """One GCL without its edge kernel = the node GEMM launches of a sub-layer (pre-projection, node_mlp.0, node_mlp.2),
    obtained as (hd_gcl_forward - hd_edge_kernel_only) back to back."""
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
    ms_gcl = _time_launches(lambda: native.check(L.hd_gcl_forward(
        cfg, native.ptr(packed), 0, 0, native.ptr(h), native.ptr(x), native.ptr(x), native.ptr(loop.sizes), B, N,
        native.ptr(ws), eid, st), "hd_gcl_forward"))
    ms_edge = _time_launches(lambda: native.check(L.hd_edge_kernel_only(
        cfg, native.ptr(packed), 0, 0, native.ptr(x), native.ptr(x), native.ptr(loop.sizes), B, N, native.ptr(ws),
        eid, st), "edge_only"))
    ms = max(ms_gcl - ms_edge, 1e-6)
    rows = float(B * N)
    flops = 2.0 * rows * (HIDDEN * 2 * HIDDEN + 2 * HIDDEN * HIDDEN + HIDDEN * HIDDEN)   # A|B, node_mlp.0, node_mlp.2
    peaks, src = measured_peaks()
    peak = 72.0 if engine == "fp32" else float(peaks.get("bf16_tflops", 1590.0))
    achieved = flops / (ms * 1e-3) / 1e12
    return {"kernel": "node GEMMs of one GCL lin::linear_tc_k (3 launches: A|B pre-projection, node_mlp.0, node_mlp.2)",
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": _profile_metrics(f"node_{engine}")[0], "peak_source": f"bf16 dense burst, {src}",
            "ms_per_sublayer": ms, "algorithmic_flops": flops,
            "note": "M = B*N = %d rows only: latency bound by construction (one wave of CTAs, K = 256..512)" % int(rows)}


def main():
    global T_STEPS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--engine", default=os.environ.get("HD_BENCH_ENGINE", "strict"), choices=["strict", "fast", "fp32"])
    ap.add_argument("--workload", default="configs1", choices=["configs1", "geom9", "sweep"])
    ap.add_argument("--steps-per-graph", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--short-gpu-reference", action="store_true",
                    help="gpu_reference: time a 50-step chain and extrapolate instead of the whole T=1000 chain")
    ap.add_argument("--cpu-steps", type=int, default=6, help="reverse steps of the CPU reference in cpu_baseline")
    ap.add_argument("--ref-steps", type=int, default=2, help="--impl reference: reverse steps per bench step")
    ap.add_argument("--no-graph", action="store_true", help="issue the loop eagerly (profiling under ncu)")
    ap.add_argument("--sizes", default="full", choices=["full", "geom"],
                    help="configs1 only - full: every molecule has N nodes (headline); geom: GEOM histogram sizes")
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
