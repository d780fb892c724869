"""The T-step reverse-diffusion loop (``DiffusionQM9.sample``, diffusion_qm9.py:347-395) as a replayed
CUDA graph.

One reverse step of the reference is ~500 ATen launches, 8 host synchronisations and one 1.6 MB
host->device copy.  Here a step is a fixed sequence of native kernels plus the two
``normal_`` launches that draw the step's noise from torch's CUDA generator in the reference's order
(``randn(B,N,3)`` then ``randn(B,N,F)``, diffusion_qm9.py:449-454).  Time and schedule scalars are read
on the device through a device-side step index, so the same captured graph is replayed for every step; the
status word (NaN guard, mask / centre-of-gravity invariants) is read once after the loop.
"""
import os

import torch

from . import native
from .noise_model import PredefinedNoiseSchedule


class ScheduleTable:
    """gamma(t) and the per-step scalars of diffusion_qm9.py:181-204,:320-334 for all T steps of a B-molecule chain.

    The reference evaluates ``gamma`` twice per step on a ``[B,1]`` tensor (``sample_p_zs_given_zt``,
    diffusion_qm9.py:314-315, called with ``s_array`` / ``t_array`` of :376-379) and every molecule uses ITS row of the
    result.  ``GammaNetwork`` is a small MLP whose rounding depends on the call shape (2.7e-4 between a
    ``[T+1,1]`` grid and ``[B,1]`` calls), so the table is built from T+1 calls of exactly that shape on the sampling
    device - ``gamma(full((B,1), k) / T)`` for k = 0..T; step s reads rows s and s+1, the final decode row 0
    (diffusion_qm9.py:296-297: ``gamma(zeros(B,1))``) - and keeps one row of scalars PER MOLECULE.

    Row k < T of ``sched`` [T+1,B,3] belongs to the k-th executed step (s = T-1-k, t = s+1); row T holds the
    final-decode scalars {alpha_0, sigma_0, sigma_x} of diffusion_qm9.py:294-304.  ``t`` [T+1] is the time fed to the
    dynamics (t/T for the steps, 0 for the final decode).

    ``refill`` recomputes the values INTO THE SAME device tensors, so CUDA graphs that captured their addresses stay
    valid when the schedule's parameters change (load_state_dict, a weight broadcast).
    """

    def __init__(self, gamma_module, T, B, device):
        self.T, self.B, self.device = T, B, torch.device(device)
        self.gamma = torch.empty(T + 1, B, dtype=torch.float32, device=device)    # gamma[k, b] = gamma(k/T) row b
        self.sched = torch.empty(T + 1, B, 3, dtype=torch.float32, device=device)
        self.t = torch.empty(T + 1, dtype=torch.float32, device=device)
        self.refill(gamma_module)

    @torch.no_grad()
    def refill(self, gamma_module):
        L, T, B, device = native.lib(), self.T, self.B, self.device
        if isinstance(gamma_module, PredefinedNoiseSchedule):
            # PredefinedNoiseSchedule is a table lookup (noise_model.py:155-160): no arithmetic, no shape dependence
            grid = (torch.arange(T + 1, device=device) / T).view(-1, 1)
            self.gamma.copy_(gamma_module(grid).reshape(-1, 1).float().expand(T + 1, B))
        else:
            for k in range(T + 1):
                # int64 fill / int -> fp32 true division, as diffusion_qm9.py:376-379
                arr = torch.full((B, 1), fill_value=k, device=device) / T
                self.gamma[k].copy_(gamma_module(arr).reshape(-1))
        order = torch.arange(T - 1, -1, -1, device=device)
        g_s = self.gamma[order].contiguous()          # [T,B]
        g_t = self.gamma[order + 1].contiguous()
        self.t[:T] = (torch.arange(T + 1, device=device) / T)[order + 1]
        self.t[T] = 0.0
        with torch.cuda.device(device):
            st = native.stream_ptr()
            native.check(L.hd_step_scalars(native.ptr(g_s), native.ptr(g_t), T * B, native.ptr(self.sched), st),
                         "hd_step_scalars")
            native.check(L.hd_final_scalars(native.ptr(self.gamma[0]), B, self.sched[T].data_ptr(), st),
                         "hd_final_scalars")
            torch.cuda.current_stream().synchronize()   # g_s / g_t are freed on return


class SamplingLoop:
    """Static buffers + captured graph for one padded batch shape (B, N).

    A reverse step is: the two ``normal_`` launches that draw the step's noise from torch's CUDA generator (the
    reference's ``randn(B,N,3)`` then ``randn(B,N,F)``, diffusion_qm9.py:449-454 - they stay torch calls so that the
    Philox stream is the reference's), then ``hd_sampler_step``: the EGNN stack and ONE kernel for everything between
    two stacks (centre-of-gravity projection, NaN guard, reverse update, step counter, the next forward's masking /
    time channel / embedding / first pre-projection).  The step index lives on the device, so one captured graph
    serves all T steps."""

    def __init__(self, model, B, N, device, steps_per_graph=8, use_graph=True):
        self.model, self.B, self.N, self.device = model, B, N, device
        self.F = model.in_node_nf
        D = 3 + self.F
        f32 = dict(dtype=torch.float32, device=device)
        self.z = torch.zeros(B, N, D, **f32)
        self.rx = torch.zeros(B, N, 3, **f32)
        self.rh = torch.zeros(B, N, self.F, **f32)
        # fix_noise (en_diffusion.py:639-642, :322-323): every draw has batch size 1 and is broadcast over the molecules
        self.fix_noise = False
        self.rx1 = torch.zeros(1, N, 3, **f32)
        self.rh1 = torch.zeros(1, N, self.F, **f32)
        self.flags = torch.zeros(1, dtype=torch.int32, device=device)
        self.sizes = torch.full((B,), N, dtype=torch.int32, device=device)
        C = model.dynamics.context_node_nf
        self.C = C
        self.context = torch.zeros(B, N, C, **f32) if C else None   # static buffer read by the captured graph
        self.x_out = torch.zeros(B, N, 3, **f32)
        self.h_out = torch.zeros(B, N, self.F, **f32)
        self.table = None
        self.use_graph = use_graph
        self.steps_per_graph = steps_per_graph
        self.ragged = False        # HD_ENGINE_RAGGED_ROWS hint of the current sizes (see ragged_rows_pay)
        self.live_rows = 0         # with the hint: sum(sizes) rounded up to whole 128-row tiles (sizes the node grids)
        self._graphs = {}          # (hint, live_rows) -> captured graph (both are baked into the captured launches)
        self._captured_ptrs = None # device addresses the captured launches embed (packed weights, schedule table)
        self.graph_steps = 0
        self.launches_per_step = None

    @property
    def graph(self):
        return self._graphs.get((self.ragged, self.live_rows, self.fix_noise))

    def _draw(self):
        """The two raw draws of a step, in the reference's order and shapes (diffusion_qm9.py:449-454)."""
        if self.fix_noise:
            self.rx1.normal_()
            self.rh1.normal_()
            self.rx.copy_(self.rx1.expand_as(self.rx))
            self.rh.copy_(self.rh1.expand_as(self.rh))
        else:
            self.rx.normal_()
            self.rh.normal_()

    @staticmethod
    def ragged_rows_pay(sizes_host, B, N):
        """Compacting the node rows pays once the padded rows B*N no longer fit one wave of node-GEMM CTAs
        (~3000 rows) and a good part of them is padding; below that it only adds set-up work."""
        return B * N >= 3072 and int(sum(int(v) for v in sizes_host)) <= 0.75 * B * N

    # -- native calls, everything enqueued on the current stream ------------------------------------
    def _common(self):
        egnn = self.model.dynamics.egnn
        engine = egnn.engine_id() | (native.ENGINE_RAGGED_ROWS if self.ragged else 0)
        return (egnn.hd_config(), native.ptr(egnn.packed_weights()), native.ptr(egnn.workspace(self.B, self.N, self.device)),
                engine, int(self.live_rows) if self.ragged else 0)

    def _begin(self):
        """Input side of the first forward from ``self.z`` (= z_T); resets the device-side step index."""
        cfg, packed, ws, engine, live = self._common()
        native.check(native.lib().hd_sampler_begin(
            cfg, packed, native.ptr(self.z), native.ptr(self.table.t), self.table.T, native.ptr(self.context), self.C,
            native.ptr(self.sizes), self.B, self.N, live, ws, native.ptr(self.flags), engine, native.stream_ptr()),
            "hd_sampler_begin")

    def _step(self):
        cfg, packed, ws, engine, live = self._common()
        self._draw()
        native.check(native.lib().hd_sampler_step(
            cfg, packed, native.ptr(self.z), native.ptr(self.rx), native.ptr(self.rh), native.ptr(self.table.t),
            native.ptr(self.table.sched), self.B, self.table.T, native.ptr(self.context), self.C,
            native.ptr(self.sizes), self.B, self.N, live, ws, native.ptr(self.flags), engine, native.stream_ptr()),
            "hd_sampler_step")

    def _final(self, norm=None):
        """``norm``: (norm_x, norm_h, bias_h) of ``unnormalize`` (diffusion_qm9.py:174-179); default: the model's."""
        m = self.model
        cfg, packed, ws, engine, live = self._common()
        self._draw()
        nx, nh, bh = norm if norm is not None else (m.norm_values[0], m.norm_values[1], m.norm_biases[1])
        native.check(native.lib().hd_sampler_final(
            cfg, packed, native.ptr(self.z), native.ptr(self.rx), native.ptr(self.rh), native.ptr(self.table.t),
            native.ptr(self.table.sched), self.B, self.table.T, native.ptr(self.context), self.C,
            native.ptr(self.sizes), self.B, self.N, live, float(nx), float(nh), float(bh), native.ptr(self.x_out),
            native.ptr(self.h_out), ws, native.ptr(self.flags), engine, native.stream_ptr()), "hd_sampler_final")

    def _capture(self, k):
        # warm-up on a side stream (also packs weights / allocates the workspace outside the capture)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        saved = self.z.clone(), self.flags.clone()
        rng = torch.cuda.get_rng_state(self.device)  # the warm-up step must not consume the sample's noise stream
        with torch.cuda.stream(s):
            self._begin()
            self._step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(k):
                self._step()
        self.z.copy_(saved[0])
        self.flags.copy_(saved[1])
        torch.cuda.synchronize(self.device)
        torch.cuda.set_rng_state(rng, self.device)
        self._graphs[(self.ragged, self.live_rows, self.fix_noise)], self.graph_steps = g, k

    def prepare(self, table):
        """Bind the schedule table and (re)capture the graph; not part of a sample's timed region."""
        self.table = table
        # captured launches hold raw device addresses: a reallocated weight image or schedule table (the model moved
        # to another device, a different configuration) invalidates every graph of this loop.  Plain weight updates
        # do not: they are repacked / refilled into the same buffers.
        ptrs = (self.model.dynamics.egnn.packed_weights().data_ptr(), table.t.data_ptr(), table.sched.data_ptr())
        if ptrs != self._captured_ptrs:
            self._graphs.clear()
            self._captured_ptrs = ptrs
        if self.use_graph and self.graph is None:
            k = self.steps_per_graph
            while k > 1 and table.T % k:
                k -= 1
            with torch.cuda.device(self.device):
                self._capture(k)

    def run(self, sizes_host, z_T=None, context=None, norm=None, fix_noise=False, on_step=None):
        """Run the whole chain; returns padded (x [B,N,3], h [B,N,F]) on the device and the status word.
        ``fix_noise``: one noise draw per step shared by all molecules.  ``on_step(s, z)``: called after the reverse step
        that produced z_s (s = T-1 ... 0) - the chain then runs step by step without the captured graph."""
        T = self.table.T
        if (context is None) != (self.context is None):
            raise ValueError("context must be given exactly when the dynamics has context_node_nf > 0")
        sizes_list = sizes_host.tolist() if hasattr(sizes_host, "tolist") else list(sizes_host)
        if len(sizes_list) != self.B or min(sizes_list) < 1 or max(sizes_list) > self.N:
            raise ValueError(f"sizes must be {self.B} values in [1, {self.N}]")
        self.fix_noise = bool(fix_noise)
        self.ragged = self.ragged_rows_pay(sizes_list, self.B, self.N)
        self.live_rows = -(-int(sum(sizes_list)) // 128) * 128 if self.ragged else 0
        with torch.cuda.device(self.device):
            # sizes first: a first capture with this hint warms up on them (live_rows is a bound on THEIR sum)
            self.sizes.copy_(torch.as_tensor(sizes_host, dtype=torch.int32), non_blocking=True)
        self.prepare(self.table)        # first chain with this hint: capture its graph
        with torch.cuda.device(self.device):
            if context is not None:
                self.context.copy_(torch.as_tensor(context, dtype=torch.float32).expand_as(self.context),
                                   non_blocking=True)
            self.flags.zero_()
            if z_T is None:
                # z_T ~ sample_combined_position_feature_noise (diffusion_qm9.py:361)
                self._draw()
                native.check(native.lib().hd_combine_noise(
                    native.ptr(self.rx), native.ptr(self.rh), native.ptr(self.sizes), self.B, self.N, self.F,
                    native.ptr(self.z), native.stream_ptr()), "hd_combine_noise")
            else:
                self.z.copy_(z_T)
            nvtx = os.environ.get("HD_NVTX") == "1"    # ranges for nsys / ncu --nvtx captures
            if nvtx:
                torch.cuda.nvtx.range_push(f"hierdiff.chain B={self.B} N={self.N} T={T}")
            self._begin()
            done = 0
            if self.graph is not None and on_step is None:
                while done + self.graph_steps <= T:
                    self.graph.replay()
                    done += self.graph_steps
            while done < T:
                self._step()
                done += 1
                if on_step is not None:
                    on_step(T - done, self.z)
            if nvtx:
                torch.cuda.nvtx.range_push("hierdiff.final_decode")
            self._final(norm)
            if nvtx:
                torch.cuda.nvtx.range_pop()
                torch.cuda.nvtx.range_pop()
        return self.x_out, self.h_out, self.flags
