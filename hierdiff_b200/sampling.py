"""The T-step reverse-diffusion loop (``DiffusionQM9.sample``, diffusion_qm9.py:347-395) as a replayed
CUDA graph.

One reverse step of the reference is ~500 ATen launches, 8 host synchronisations and one 1.6 MB
host->device copy.  Here a step is a fixed sequence of native kernels plus the two
``normal_`` launches that draw the step's noise from torch's CUDA generator in the reference's order
(``randn(B,N,3)`` then ``randn(B,N,F)``, diffusion_qm9.py:449-454).  Time and schedule scalars are read
on the device through a step counter, so the same captured graph is replayed for every step; the
status word (NaN guard, mask / centre-of-gravity invariants) is read once after the loop.
"""
import os

import torch

from . import native


class ScheduleTable:
    """gamma(t) and the per-step scalars of diffusion_qm9.py:181-204,:320-334 for all T steps.

    Row k < T belongs to the k-th executed step (s = T-1-k, t = s+1); row T holds the final-decode
    scalars {alpha_0, sigma_0, sigma_x} of diffusion_qm9.py:294-304 and time 0.
    """

    def __init__(self, gamma_module, T, device):
        L = native.lib()
        self.T = T
        with torch.no_grad():
            # the reference evaluates gamma on s/T and (s+1)/T, int64 / int -> fp32 true division (:376-379)
            grid = (torch.arange(T + 1, device=device) / T).view(-1, 1)
            gamma = gamma_module(grid).reshape(-1).float().contiguous()   # gamma[k] = gamma(k/T)
        self.gamma = gamma
        order = torch.arange(T - 1, -1, -1, device=device)
        g_s = gamma[order].contiguous()
        g_t = gamma[order + 1].contiguous()
        self.sched = torch.empty(T + 1, 3, dtype=torch.float32, device=device)
        self.t = torch.empty(T + 1, dtype=torch.float32, device=device)
        self.t[:T] = grid.view(-1)[order + 1]
        self.t[T] = 0.0
        with torch.cuda.device(device):
            st = native.stream_ptr()
            native.check(L.hd_step_scalars(native.ptr(g_s), native.ptr(g_t), T, native.ptr(self.sched), st),
                         "hd_step_scalars")
            native.check(L.hd_final_scalars(native.ptr(gamma[:1].contiguous()), 1,
                                            self.sched[T].data_ptr(), st), "hd_final_scalars")
            torch.cuda.current_stream().synchronize()


class SamplingLoop:
    """Static buffers + captured graph for one padded batch shape (B, N)."""

    def __init__(self, model, B, N, device, steps_per_graph=8, use_graph=True):
        self.model, self.B, self.N, self.device = model, B, N, device
        self.F = model.in_node_nf
        D = 3 + self.F
        f32 = dict(dtype=torch.float32, device=device)
        self.z = torch.zeros(B, N, D, **f32)
        self.eps = torch.zeros(B, N, D, **f32)
        self.rx = torch.zeros(B, N, 3, **f32)
        self.rh = torch.zeros(B, N, self.F, **f32)
        self.t_cur = torch.zeros(B, **f32)
        self.sched_cur = torch.zeros(3, **f32)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.flags = torch.zeros(1, dtype=torch.int32, device=device)
        self.sizes = torch.full((B,), N, dtype=torch.int32, device=device)
        C = model.dynamics.context_node_nf
        self.context = torch.zeros(B, N, C, **f32) if C else None   # static buffer read by the captured graph
        self.x_out = torch.zeros(B, N, 3, **f32)
        self.h_out = torch.zeros(B, N, self.F, **f32)
        self.table = None
        self.use_graph = use_graph
        self.steps_per_graph = steps_per_graph
        self.ragged = False        # HD_ENGINE_RAGGED_ROWS hint of the current sizes (see ragged_rows_pay)
        self.live_rows = 0         # with the hint: sum(sizes) rounded up to whole 128-row tiles (sizes the node grids)
        self._graphs = {}          # (hint, live_rows) -> captured graph (both are baked into the captured launches)
        self.graph_steps = 0
        self.launches_per_step = None

    @property
    def graph(self):
        return self._graphs.get((self.ragged, self.live_rows))

    @staticmethod
    def ragged_rows_pay(sizes_host, B, N):
        """Compacting the node rows pays once the padded rows B*N no longer fit one wave of node-GEMM CTAs
        (~3000 rows) and a good part of them is padding; below that it only adds set-up work."""
        return B * N >= 3072 and int(sum(int(v) for v in sizes_host)) <= 0.75 * B * N

    # -- one reverse step, everything enqueued on the current stream --------------------------------
    def _step(self):
        L, m = native.lib(), self.model
        st = native.stream_ptr()
        native.check(L.hd_loop_fetch(native.ptr(self.counter), native.ptr(self.table.t), native.ptr(self.table.sched),
                                     self.B, native.ptr(self.t_cur), native.ptr(self.sched_cur), st), "hd_loop_fetch")
        m.dynamics.forward_sizes(self.t_cur, self.z, self.sizes, flags=self.flags, out=self.eps, context=self.context,
                                 ragged=self.ragged, live_rows=self.live_rows)
        self.rx.normal_()
        self.rh.normal_()
        native.check(L.hd_reverse_step(native.ptr(self.z), native.ptr(self.eps), native.ptr(self.rx),
                                       native.ptr(self.rh), native.ptr(self.sizes), self.B, self.N, self.F,
                                       native.ptr(self.sched_cur), 0, native.ptr(self.z), native.ptr(self.flags),
                                       st), "hd_reverse_step")

    def _final(self):
        L, m = native.lib(), self.model
        st = native.stream_ptr()
        native.check(L.hd_loop_fetch(native.ptr(self.counter), native.ptr(self.table.t), native.ptr(self.table.sched),
                                     self.B, native.ptr(self.t_cur), native.ptr(self.sched_cur), st), "hd_loop_fetch")
        m.dynamics.forward_sizes(self.t_cur, self.z, self.sizes, flags=self.flags, out=self.eps, context=self.context,
                                 ragged=self.ragged, live_rows=self.live_rows)
        self.rx.normal_()
        self.rh.normal_()
        nv, nb = m.norm_values, m.norm_biases
        native.check(L.hd_final_decode(native.ptr(self.z), native.ptr(self.eps), native.ptr(self.rx),
                                       native.ptr(self.rh), native.ptr(self.sizes), self.B, self.N, self.F,
                                       native.ptr(self.sched_cur), 0, float(nv[0]), float(nv[1]), float(nb[1]),
                                       native.ptr(self.x_out), native.ptr(self.h_out), st), "hd_final_decode")

    def _capture(self, k):
        # warm-up on a side stream (also packs weights / allocates the workspace outside the capture)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        saved = self.z.clone(), self.counter.clone(), self.flags.clone()
        rng = torch.cuda.get_rng_state(self.device)  # the warm-up step must not consume the sample's noise stream
        with torch.cuda.stream(s):
            self.counter.zero_()
            self._step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(k):
                self._step()
        self.z.copy_(saved[0])
        self.counter.copy_(saved[1])
        self.flags.copy_(saved[2])
        torch.cuda.synchronize(self.device)
        torch.cuda.set_rng_state(rng, self.device)
        self._graphs[(self.ragged, self.live_rows)], self.graph_steps = g, k

    def prepare(self, table):
        """Bind the schedule table and (re)capture the graph; not part of a sample's timed region."""
        self.table = table
        if self.use_graph and self.graph is None:
            k = self.steps_per_graph
            while k > 1 and table.T % k:
                k -= 1
            with torch.cuda.device(self.device):
                self._capture(k)

    def run(self, sizes_host, z_T=None, context=None):
        """Run the whole chain; returns padded (x [B,N,3], h [B,N,F]) on the device and the status word."""
        T = self.table.T
        if (context is None) != (self.context is None):
            raise ValueError("context must be given exactly when the dynamics has context_node_nf > 0")
        sizes_list = sizes_host.tolist() if hasattr(sizes_host, "tolist") else list(sizes_host)
        self.ragged = self.ragged_rows_pay(sizes_list, self.B, self.N)
        self.live_rows = -(-int(sum(sizes_list)) // 128) * 128 if self.ragged else 0
        self.prepare(self.table)        # first chain with this hint: capture its graph
        with torch.cuda.device(self.device):
            self.sizes.copy_(torch.as_tensor(sizes_host, dtype=torch.int32), non_blocking=True)
            if context is not None:
                self.context.copy_(torch.as_tensor(context, dtype=torch.float32).expand_as(self.context),
                                   non_blocking=True)
            self.counter.zero_()
            self.flags.zero_()
            if z_T is None:
                # z_T ~ sample_combined_position_feature_noise (diffusion_qm9.py:361)
                self.rx.normal_()
                self.rh.normal_()
                native.check(native.lib().hd_combine_noise(
                    native.ptr(self.rx), native.ptr(self.rh), native.ptr(self.sizes), self.B, self.N, self.F,
                    native.ptr(self.z), native.stream_ptr()), "hd_combine_noise")
            else:
                self.z.copy_(z_T)
            nvtx = os.environ.get("HD_NVTX") == "1"    # ranges for nsys / ncu --nvtx captures
            if nvtx:
                torch.cuda.nvtx.range_push(f"hierdiff.chain B={self.B} N={self.N} T={T}")
            done = 0
            if self.graph is not None:
                while done + self.graph_steps <= T:
                    self.graph.replay()
                    done += self.graph_steps
            while done < T:
                self._step()
                done += 1
            if nvtx:
                torch.cuda.nvtx.range_push("hierdiff.final_decode")
            self._final()
            if nvtx:
                torch.cuda.nvtx.range_pop()
                torch.cuda.nvtx.range_pop()
        return self.x_out, self.h_out, self.flags
