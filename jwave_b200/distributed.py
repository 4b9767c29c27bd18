"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

Two ways the path spreads over the GPUs of a box:

* independent signals / images (configs 2-4): `shard_range` splits the batch into contiguous
  blocks, every rank transforms its own block, NO collective touches the data;
* one large volume (config 5): `SlabVolumeTransform` keeps the volume slab-decomposed along the
  outer index i.  The two inner axes (k, j) are transformed locally on the owned slices; the
  outer axis needs every i for a given (j, k), so one all-to-all re-slabs the volume along j, the
  i-pass runs locally, and a second all-to-all restores the i-slab layout - the layout
  BasicTransform.forward(double[][][]) returns (BasicTransform.java:509-566), just distributed.

The local compute is injected (`axis_fn`), so the exchange logic is testable on CPU ranks with
the `gloo` backend; on GPUs it is DeviceTransforms.axis (libjwave_cuda.so) and the collective
runs over NCCL / NVLink.
"""
import os

import torch
import torch.distributed as dist

FORWARD, REVERSE = 0, 1
FWT, WPT = 0, 1


def shard_range(total, rank, world):
    """Contiguous block [start, stop) of `total` independent items owned by `rank`."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class SlabVolumeTransform:
    """3-D FWT / WPT of a P x Q x R volume held as i-slabs: rank g owns [g P/W, (g+1) P/W) x Q x R.

    axis_fn(kind, direction, x, outer, n, inner, level) -> tensor like x: the 1-D transform along
    the middle axis of the dense [outer][n][inner] view of x (jwc_axis_dev semantics).
    """

    def __init__(self, axis_fn, kind=FWT, group=None):
        self.axis_fn = axis_fn
        self.kind = kind
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # -- exchanges ---------------------------------------------------------------------------------
    def _to_j_slabs(self, x):
        """[P/W][Q][R] on every rank -> [P][Q/W][R] on every rank (one all-to-all)."""
        W = self.world
        if W == 1:
            return x
        p, Q, R = x.shape
        send = x.view(p, W, Q // W, R).permute(1, 0, 2, 3).contiguous()  # [dest][i][j_local][k]
        recv = torch.empty_like(send)                                       # [src][i_of_src][j_local][k]
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.view(W * p, Q // W, R)  # src-major order is i order

    def _to_i_slabs(self, y, out=None):
        """[P][Q/W][R] on every rank -> [P/W][Q][R] on every rank (one all-to-all); the unpack copy
        lands directly in `out` when one is given."""
        W = self.world
        if W == 1:
            if out is None:
                return y
            out.copy_(y.view_as(out))
            return out
        P, q, R = y.shape
        send = y.contiguous()               # chunk d = rows of rank d's i range
        recv = torch.empty_like(send)       # [src][i_local][j_of_src][k]
        dist.all_to_all_single(recv, send, group=self.group)
        if out is None:
            out = torch.empty(P // W, W * q, R, dtype=y.dtype, device=y.device)
        out.view(P // W, W, q, R).copy_(recv.view(W, P // W, q, R).permute(1, 0, 2, 3))
        return out

    # -- transforms --------------------------------------------------------------------------------
    def _check(self, x, P):
        if x.dim() != 3 or x.shape[0] * self.world != P:
            raise ValueError("expected this rank's [P/W][Q][R] slab")
        if x.shape[1] % self.world:
            raise ValueError("Q must be divisible by the number of ranks")

    def forward(self, slab, P, lvlP, lvlQ, lvlR, out=None):
        """BasicTransform.java:509-566 with its level shift (F5): axis k gets lvlQ, axis j gets
        lvlP, then axis i gets lvlR."""
        self._check(slab, P)
        p, Q, R = slab.shape
        t = self.axis_fn(self.kind, FORWARD, slab, p * Q, R, 1, lvlQ)   # rows of every slice
        t = self.axis_fn(self.kind, FORWARD, t, p, Q, R, lvlP)          # columns of every slice
        y = self._to_j_slabs(t)
        y = self.axis_fn(self.kind, FORWARD, y, 1, P, y.shape[1] * R, lvlR)
        return self._to_i_slabs(y, out)

    def reverse(self, slab, P, lvlP, lvlQ, lvlR, out=None):
        """BasicTransform.java:602-659: 2-D reverse of every slice (columns, then rows), then axis i."""
        self._check(slab, P)
        p, Q, R = slab.shape
        t = self.axis_fn(self.kind, REVERSE, slab, p, Q, R, lvlP)
        t = self.axis_fn(self.kind, REVERSE, t, p * Q, R, 1, lvlQ)
        y = self._to_j_slabs(t)
        y = self.axis_fn(self.kind, REVERSE, y, 1, P, y.shape[1] * R, lvlR)
        return self._to_i_slabs(y, out)

    def exchange_bytes(self, slab):
        """Bytes this rank sends per all-to-all (the NVLink term of the roofline)."""
        return slab.numel() * slab.element_size() * (self.world - 1) // max(self.world, 1)


class PeerSlabVolumeTransform:
    """The same slab-decomposed 3-D FWT with the exchanges folded into the kernels: every rank maps the
    slabs of its peers (torch symmetric memory over NVLink / NVSwitch) and the axis pass that precedes an
    exchange stores its output rows straight into the peer that owns them (jwc_axis_dev_remote), so there
    is no all-to-all, no pack and no unpack - three device-side barriers per direction instead.

    forward / reverse return a view of an internal symmetric buffer in the i-slab layout [P/W][Q][R]; it
    stays valid until the next-but-one call.  FWT only (the WPT has no fused strided kernels), P/W and
    Q/W powers of two, R a multiple of 8; use SlabVolumeTransform otherwise."""

    def __init__(self, dev, P, Q, R, group=None, exchange="stores", chunks=None):
        """exchange = "stores": the axis kernels store into the peers (fewest passes over the data, but a
        strided-axis CTA owns 64-byte pieces of each row - fine for 2 peers, slow across 8);
        "copies": the passes stay local and W - 1 strided device copies per exchange write the blocks
        straight into the peers' slabs in their final layout (wide rows; still no pack, no unpack, no
        NCCL)."""
        import torch.distributed._symmetric_memory as symm
        if exchange not in ("stores", "copies"):
            raise ValueError("exchange must be 'stores' or 'copies'")
        self.exchange = exchange
        self.dev, self.group = dev, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        W = self.world
        self.P, self.Q, self.R = P, Q, R
        self.p, self.q = P // W, Q // W
        ok = (P % W == 0 and Q % W == 0 and self.p & (self.p - 1) == 0 and self.q & (self.q - 1) == 0
              and R % 8 == 0 and 2 <= W <= 8)
        if not ok:
            raise ValueError("PeerSlabVolumeTransform: P/W and Q/W must be powers of two, R % 8 == 0, 2 <= W <= 8")
        name = (group or dist.group.WORLD).group_name
        device = dev.device
        self._bufs, self._hdl, self._ptrs = {}, {}, {}
        for key, numel in (("J", P * self.q * R), ("I0", self.p * Q * R), ("I1", self.p * Q * R)):
            t = symm.empty(numel, dtype=torch.float64, device=device)
            h = symm.rendezvous(t, name)
            self._bufs[key], self._hdl[key] = t, h
            self._peer = getattr(self, "_peer", {})
            self._peer[key] = [h.get_buffer(r, (numel,), torch.float64) for r in range(W)]
            self._ptrs[key] = [t_.data_ptr() for t_ in self._peer[key]]
        # "copies": chunks per re-cut (JWB_SLAB_CHUNKS overrides); each chunk keeps whole 16-column blocks.  Measured
        # on 1024^3 (gpurun_out -> profiles/r02_bench_c5_n*.json): 2 GPUs 84 / 92 / 88 GS/s with 2 / 4 / 8 chunks,
        # 4 GPUs 177 / 175, 8 GPUs 297 / 263 / 225 - the slabs of 8 GPUs are 128 rows, finer chunks stop paying
        C = int(os.environ.get("JWB_SLAB_CHUNKS", chunks or (2 if W >= 8 else 4)))
        while C > 1 and (self.p % C or self.q % C or (self.q // C) * R % 16):
            C //= 2
        self.chunks = max(C, 1)
        # several copy streams: one stream keeps one copy engine busy (~400 GB/s over NVLink measured), the re-cut
        # of a chunk is W * C independent 2-D copies
        ns = int(os.environ.get("JWB_SLAB_COPY_STREAMS", 4))
        self._copy_streams = [torch.cuda.Stream(device=device, priority=-1) for _ in range(max(ns, 1))]
        self._tmp = torch.empty(self.p * Q * R, dtype=torch.float64, device=device)
        self._tmp2 = torch.empty(self.p * Q * R, dtype=torch.float64, device=device) if exchange == "copies" else None
        self._flip = 0

    @staticmethod
    def _lg(v):
        return v.bit_length() - 1

    def _run(self, direction, slab, lvlP, lvlQ, lvlR):
        p, q, P, Q, R, rank = self.p, self.q, self.P, self.Q, self.R, self.rank
        if tuple(slab.shape) != (p, Q, R):
            raise ValueError("expected this rank's [P/W][Q][R] slab")
        dev, J = self.dev, self._bufs["J"]
        key = "I%d" % self._flip
        self._flip ^= 1
        if self.exchange == "copies":
            return self._run_copies(direction, slab, lvlP, lvlQ, lvlR, key)
        if direction == FORWARD:   # BasicTransform.java:509-566 (F5): k gets lvlQ, j gets lvlP, i gets lvlR
            dev.axis(FWT, FORWARD, slab, p * Q, R, 1, lvlQ, out=self._tmp.view(p, Q, R))
            self._hdl["J"].barrier()
            dev.axis_remote(FWT, FORWARD, self._tmp, p, Q, R, lvlP, self._ptrs["J"], 1, self._lg(q),
                            outer_stride=q * R, row_stride=R, base_off=rank * p * q * R)
        else:                      # BasicTransform.java:602-659: columns, then rows, then axis i
            dev.axis(FWT, REVERSE, slab, p, Q, R, lvlP, out=self._tmp.view(p, Q, R))
            self._hdl["J"].barrier()
            dev.axis_remote(FWT, REVERSE, self._tmp, p * Q, R, 1, lvlQ, self._ptrs["J"], 2, self._lg(q),
                            lg_hi=self._lg(Q), outer_stride=q * R, row_stride=R, base_off=rank * p * q * R)
        self._hdl["J"].barrier()   # every rank's rows have landed in my j-slab
        dev.axis_remote(FWT, direction, J, 1, P, q * R, lvlR, self._ptrs[key], 1, self._lg(p),
                        row_stride=Q * R, base_off=rank * q * R)
        self._hdl[key].barrier()   # ... and in my i-slab
        return self._bufs[key].view(p, Q, R)

    # ---- "copies": chunked re-cuts on the copy engines, overlapped with the local passes -------------------
    # Layouts.  i-slab: [p][Q][R] (rank g owns i in [g p, (g+1) p)) - the layout BasicTransform returns, just
    # distributed.  j-slab: my j range [g q, (g+1) q) for ALL i, held as C dense sub-slabs [C][P][q/C][R] so that
    # every sub-slab is one dense [outer = 1][n = P][inner = q/C * R] problem for the i pass.
    def _local_ij(self, direction, src, sl, lvlP, lvlQ):
        """the two local passes (axes k and j) of the slices `sl` of an i-slab: src[sl] -> self._b[sl]"""
        dev, Q, R = self.dev, self.Q, self.R
        a, b = self._tmp.view(self.p, Q, R), self._tmp2.view(self.p, Q, R)
        S = sl.stop - sl.start
        if direction == FORWARD:   # BasicTransform.java:509-566 (F5): k gets lvlQ, j gets lvlP
            dev.axis(FWT, FORWARD, src[sl], S * Q, R, 1, lvlQ, out=a[sl])
            dev.axis(FWT, FORWARD, a[sl], S, Q, R, lvlP, out=b[sl])
        else:                      # BasicTransform.java:602-659: columns, then rows
            dev.axis(FWT, REVERSE, src[sl], S, Q, R, lvlP, out=a[sl])
            dev.axis(FWT, REVERSE, a[sl], S * Q, R, 1, lvlQ, out=b[sl])

    def _send_to_j(self, c, comp, cs):
        """slices chunk c of my i-slab result (self._tmp2) -> the j-slabs of all ranks (DMA, copy stream)"""
        p, q, P, Q, R, rank, W, C = self.p, self.q, self.P, self.Q, self.R, self.rank, self.world, self.chunks
        S, qc = p // C, q // C
        ev = torch.cuda.Event()
        ev.record(comp)
        for st in cs:
            st.wait_event(ev)
        row = qc * R * 8                                   # one (slice, sub-slab) row: qc * R contiguous doubles
        src0 = self._tmp2.data_ptr() + c * S * Q * R * 8   # b[c S][0][0]
        n = 0
        for k in range(W):  # start with my own block, then round the ring so the peers are hit evenly
            d = (rank + k) % W
            for cc in range(C):
                # dst: J_d[cc][rank p + c S + s][:][:], s = 0 .. S - 1 (dense rows); src: b[c S + s][d q + cc qc][:]
                dst = self._ptrs["J"][d] + ((cc * P + rank * p + c * S) * qc * R) * 8
                src = src0 + ((d * q + cc * qc) * R) * 8
                self.dev.copy2d(dst, row, src, Q * R * 8, row, S, stream=cs[n % len(cs)])
                n += 1

    def _send_to_i(self, c, key, comp, cs):
        """sub-slab c of my j-slab result (self._y) -> the i-slabs `key` of all ranks, in their final layout"""
        p, q, P, Q, R, rank, W, C = self.p, self.q, self.P, self.Q, self.R, self.rank, self.world, self.chunks
        qc = q // C
        ev = torch.cuda.Event()
        ev.record(comp)
        for st in cs:
            st.wait_event(ev)
        row = qc * R * 8
        for k in range(W):
            d = (rank + k) % W
            # dst: I_d[i][rank q + c qc][:], i = 0 .. p - 1 (pitch Q R); src: y[c][d p + i][:][:] (dense rows)
            dst = self._ptrs[key][d] + ((rank * q + c * qc) * R) * 8
            src = self._y.data_ptr() + ((c * P + d * p) * qc * R) * 8
            self.dev.copy2d(dst, Q * R * 8, src, row, row, p, stream=cs[(c * W + k) % len(cs)])

    def _run_copies(self, direction, slab, lvlP, lvlQ, lvlR, key, compute=True, copies=True):
        """The reference's layout on both sides (i-slabs in, i-slabs out): two re-cuts per direction, each cut into
        C chunks whose copies run on the copy engines (copy streams) beside the next chunk's axis passes; only the
        last chunk of a re-cut is exposed.  The j-slab is held as C dense sub-slabs [C][P][q/C][R] so that the
        second re-cut can follow the i pass sub-slab by sub-slab.  `compute` / `copies`: measure()."""
        p, q, P, Q, R, C = self.p, self.q, self.P, self.Q, self.R, self.chunks
        dev = self.dev
        comp, cs = torch.cuda.current_stream(dev.device), self._copy_streams
        S, qc = p // C, q // C
        Jc = self._bufs["J"].view(C, P, qc, R)

        def join_copies():
            for st in cs:
                comp.wait_stream(st)
        # No barrier on entry: every rank passed the closing barrier of the previous call only after its own i
        # passes had read its j-slab, and after the local passes of that call had read the i-slab it was given.
        for c in range(C):                               # re-cut 1: i-slabs -> j-slabs, behind the k / j passes
            if compute:
                self._local_ij(direction, slab, slice(c * S, (c + 1) * S), lvlP, lvlQ)
            if copies:
                self._send_to_j(c, comp, cs)
        join_copies()
        self._hdl["J"].barrier()                         # every rank's rows have landed in my j-slab
        self._y = self._tmp.view(C, P, qc, R)
        for c in range(C):                               # the i pass, sub-slab by sub-slab
            if compute:
                dev.axis(FWT, direction, Jc[c], 1, P, qc * R, lvlR, out=self._y[c])
            if copies:                                   # re-cut 2: j-slabs -> i-slabs
                self._send_to_i(c, key, comp, cs)
        join_copies()
        self._hdl[key].barrier()
        return self._bufs[key].view(p, Q, R)

    # ---- coefficients left in the j-slab layout: ONE re-cut per direction ---------------------------------
    # Plain layouts on both sides (i-slab [p][Q][R], j-slab [P][q][R]), so every 2-D copy of a re-cut moves rows of
    # q * R contiguous doubles (1 MiB for 1024^3 on 8 GPUs): the copy engines' rate over NVLink falls with the row
    # width (715 GB/s at 1 MiB rows, 380 GB/s at 256 KiB - profiles/r02_slab_recut.md), which is what the
    # sub-slab form above pays for being able to overlap its second re-cut.
    def _copy_rows_to_j(self, c, cs):
        """slices chunk c of my k/j-pass result (self._tmp2, i-slab) -> rows [rank p + c S, + S) of every j-slab"""
        p, q, Q, R, rank, W = self.p, self.q, self.Q, self.R, self.rank, self.world
        S = p // self.chunks
        for k in range(W):  # my own block first, then round the ring so the peers are hit evenly
            d = (rank + k) % W
            dst = self._ptrs["J"][d] + ((rank * p + c * S) * q * R) * 8
            src = self._tmp2.data_ptr() + ((c * S) * Q * R + d * q * R) * 8
            self.dev.copy2d(dst, q * R * 8, src, Q * R * 8, q * R * 8, S, stream=cs[k % len(cs)])

    def _copy_rows_to_i(self, c, key, cs):
        """rows of slices chunk c of every rank's i range, from my i-pass result (self._tmp2, j-slab) -> their i-slabs"""
        p, q, Q, R, rank, W = self.p, self.q, self.Q, self.R, self.rank, self.world
        S = p // self.chunks
        for k in range(W):
            d = (rank + k) % W
            dst = self._ptrs[key][d] + ((c * S) * Q * R + rank * q * R) * 8
            src = self._tmp2.data_ptr() + ((d * p + c * S) * q * R) * 8
            self.dev.copy2d(dst, Q * R * 8, src, q * R * 8, q * R * 8, S, stream=cs[k % len(cs)])

    def _forward_t(self, slab, lvlP, lvlQ, lvlR, compute=True, copies=True):
        p, q, P, Q, R, C = self.p, self.q, self.P, self.Q, self.R, self.chunks
        comp, cs = torch.cuda.current_stream(self.dev.device), self._copy_streams
        S = p // C
        for c in range(C):                               # k and j passes of chunk c; its rows leave behind them
            if compute:
                self._local_ij(FORWARD, slab, slice(c * S, (c + 1) * S), lvlP, lvlQ)
            if copies:
                ev = torch.cuda.Event()
                ev.record(comp)
                for st in cs:
                    st.wait_event(ev)
                self._copy_rows_to_j(c, cs)
        for st in cs:
            comp.wait_stream(st)
        self._hdl["J"].barrier()                         # every rank's rows have landed in my j-slab
        y = self._tmp.view(P, q, R)
        if compute:
            self.dev.axis(FWT, FORWARD, self._bufs["J"].view(P, q, R), 1, P, q * R, lvlR, out=y)
        self._hdl["J"].barrier()                         # closing barrier: my j-slab may be refilled by the next call
        return y

    def _reverse_t(self, coef, lvlP, lvlQ, lvlR, key, compute=True, copies=True):
        p, q, P, Q, R, C = self.p, self.q, self.P, self.Q, self.R, self.chunks
        comp, cs = torch.cuda.current_stream(self.dev.device), self._copy_streams
        S = p // C
        other = "I1" if key == "I0" else "I0"
        I, out = self._bufs[key].view(p, Q, R), self._bufs[other].view(p, Q, R)
        a = self._tmp.view(p, Q, R)
        if compute:                                      # axis i first (ParallelTransform.java:193)
            self.dev.axis(FWT, REVERSE, coef, 1, P, q * R, lvlR, out=self._tmp2.view(P, q, R))
        ev = torch.cuda.Event()
        ev.record(comp)
        landed = []
        for st in cs:
            st.wait_event(ev)
        for c in range(C):                               # the re-cut runs AHEAD of the slices' passes, chunk by chunk
            if copies:
                self._copy_rows_to_i(c, key, cs)
            evs = []
            for st in cs:
                e = torch.cuda.Event()
                e.record(st)
                evs.append(e)
            landed.append(evs)
        for c in range(C):
            for e in landed[c]:
                comp.wait_event(e)
            self._hdl[key].barrier()                     # chunk c of every rank has landed in my i-slab
            if compute:                                  # BasicTransform.java:602-659 per slice: columns, then rows
                sl = slice(c * S, (c + 1) * S)
                self.dev.axis(FWT, REVERSE, I[sl], S, Q, R, lvlP, out=a[sl])
                self.dev.axis(FWT, REVERSE, a[sl], S * Q, R, 1, lvlQ, out=out[sl])
        return out

    def forward_t(self, slab, P, lvlP, lvlQ, lvlR):
        """Forward transform that leaves the coefficients in the j-slab layout [P][q][R] (rank g owns
        j in [g q, (g+1) q)): one re-cut instead of two.  The result lives in an internal
        buffer: valid until the next call."""
        if self.exchange != "copies" or tuple(slab.shape) != (self.p, self.Q, self.R):
            raise ValueError("forward_t: exchange='copies' and this rank's [P/W][Q][R] slab")
        return self._forward_t(slab, lvlP, lvlQ, lvlR)

    def reverse_t(self, coef_t, P, lvlP, lvlQ, lvlR):
        """Reverse of forward_t: j-slab coefficients [P][q][R] in, i-slab of samples out (an internal buffer, valid
        until the next-but-one call; the next call may take it as its input).  Axis i is rebuilt first - the order of
        ParallelTransform.reverse (ParallelTransform.java:193), rounding-level different from BasicTransform's."""
        if self.exchange != "copies" or tuple(coef_t.shape) != (self.P, self.q, self.R):
            raise ValueError("reverse_t: exchange='copies' and this rank's [P][Q/W][R] coefficients")
        self._flip ^= 1
        return self._reverse_t(coef_t, lvlP, lvlQ, lvlR, "I%d" % self._flip)

    @staticmethod
    def t_to_dense(coef_t):
        """the j-slab coefficients as a dense [P][q][R] array (they already are)"""
        return coef_t

    def exchange_bytes(self, slab=None):
        """Bytes this rank sends to OTHER GPUs per re-cut (the NVLink term of the step)."""
        return self.p * self.Q * self.R * 8 * (self.world - 1) // self.world

    def measure(self, slab, lvl, reps=3, coef="i"):
        """Device time (ms, max over ranks is the caller's business) of one forward call as run, with the local
        passes only and with the copies only - what the re-cuts cost alone and how much of that the overlap hides."""
        if self.exchange != "copies":
            return {}
        out = {}

        def call(i, **kw):
            if coef == "j":
                return self._forward_t(slab, lvl, lvl, lvl, **kw)
            return self._run_copies(FORWARD, slab, lvl, lvl, lvl, "I%d" % (i & 1), **kw)
        for tag, kw in (("full", {}), ("compute_only", {"copies": False}), ("copies_only", {"compute": False})):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            call(0, **kw)
            torch.cuda.synchronize()
            evs[0].record()
            for i in range(reps):
                call(i, **kw)
            evs[1].record()
            torch.cuda.synchronize()
            out[tag + "_ms"] = evs[0].elapsed_time(evs[1]) / reps
        return out

    def forward(self, slab, P, lvlP, lvlQ, lvlR, out=None):
        y = self._run(FORWARD, slab, lvlP, lvlQ, lvlR)
        return y if out is None else out.copy_(y)

    def reverse(self, slab, P, lvlP, lvlQ, lvlR, out=None):
        y = self._run(REVERSE, slab, lvlP, lvlQ, lvlR)
        return y if out is None else out.copy_(y)


def device_axis_fn(dev):
    """axis_fn backed by libjwave_cuda.so through jwave_b200.device.DeviceTransforms."""
    def fn(kind, direction, x, outer, n, inner, level):
        return dev.axis(kind, direction, x, outer, n, inner, level)
    return fn
