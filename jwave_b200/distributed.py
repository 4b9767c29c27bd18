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
        # "copies": chunks per re-cut (JWB_SLAB_CHUNKS overrides); each chunk keeps whole 16-column blocks
        import os
        C = int(os.environ.get("JWB_SLAB_CHUNKS", chunks or 4))
        while C > 1 and (self.p % C or self.q % C or (self.q // C) * R % 16):
            C //= 2
        self.chunks = max(C, 1)
        self._copy_stream = torch.cuda.Stream(device=device)
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

    def _run_copies(self, direction, slab, lvlP, lvlQ, lvlR, key, compute=True, copies=True):
        """Chunked, overlapped form.  Phase A: the owned slices are cut into C chunks; chunk c runs its two local
        passes (k and j) on the compute stream while the copy stream writes chunk c - 1 straight into the peers'
        j-slabs.  Phase B: the j-slab is held as C dense sub-slabs [P][q/C][R]; sub-slab c runs the i pass while
        the copy stream writes sub-slab c - 1 into the peers' i-slabs in their final layout.  Only the last
        chunk of each re-cut is exposed.  `compute` / `copies` switch either half off (measure())."""
        p, q, P, Q, R, rank, W, C = self.p, self.q, self.P, self.Q, self.R, self.rank, self.world, self.chunks
        dev, J = self.dev, self._bufs["J"]
        comp, cs = torch.cuda.current_stream(dev.device), self._copy_stream
        a, b = self._tmp.view(p, Q, R), self._tmp2.view(p, Q, R)
        S, qc = p // C, q // C
        Jc = J.view(C, P, qc, R)                      # my j-slab as C dense sub-slabs
        peerJ = [t.view(C, P, qc, R) for t in self._peer["J"]]
        peerI = [t.view(p, W, C, qc, R) for t in self._peer[key]]
        y = self._tmp.view(C, P, qc, R)               # phase B output (phase A's `a` is dead by then)
        # No barrier on entry: every rank passed the closing barrier of the previous call only after its own i
        # passes had read its j-slab, and after phase A of that call had read the i-slab it was given.
        for c in range(C):
            sl = slice(c * S, (c + 1) * S)
            if compute:
                if direction == FORWARD:   # BasicTransform.java:509-566 (F5): k gets lvlQ, j gets lvlP
                    dev.axis(FWT, FORWARD, slab[sl], S * Q, R, 1, lvlQ, out=a[sl])
                    dev.axis(FWT, FORWARD, a[sl], S, Q, R, lvlP, out=b[sl])
                else:                      # BasicTransform.java:602-659: columns, then rows
                    dev.axis(FWT, REVERSE, slab[sl], S, Q, R, lvlP, out=a[sl])
                    dev.axis(FWT, REVERSE, a[sl], S * Q, R, 1, lvlQ, out=b[sl])
            if copies:
                ev = torch.cuda.Event()
                ev.record(comp)
                with torch.cuda.stream(cs):
                    cs.wait_event(ev)
                    src = b[sl].view(S, W, C, qc, R)
                    for k in range(W):  # start with my own block, then round the ring so the peers are hit evenly
                        d = (rank + k) % W
                        peerJ[d][:, rank * p + c * S:rank * p + (c + 1) * S].copy_(src[:, d].permute(1, 0, 2, 3))
        comp.wait_stream(cs)
        self._hdl["J"].barrier()                      # every rank's rows have landed in my j-slab
        for c in range(C):
            if compute:
                dev.axis(FWT, direction, Jc[c], 1, P, qc * R, lvlR, out=y[c])
            if copies:
                ev = torch.cuda.Event()
                ev.record(comp)
                with torch.cuda.stream(cs):
                    cs.wait_event(ev)
                    for k in range(W):
                        d = (rank + k) % W
                        peerI[d][:, rank, c].copy_(y[c].view(W, p, qc, R)[d])
        comp.wait_stream(cs)
        self._hdl[key].barrier()
        return self._bufs[key].view(p, Q, R)

    def exchange_bytes(self, slab=None):
        """Bytes this rank sends to OTHER GPUs per re-cut (the NVLink term of the step)."""
        return self.p * self.Q * self.R * 8 * (self.world - 1) // self.world

    def measure(self, slab, lvl, reps=3):
        """Device time (ms, max over ranks is the caller's business) of one forward call as run, with the local
        passes only and with the copies only - what the re-cuts cost alone and how much of that the overlap hides."""
        if self.exchange != "copies":
            return {}
        out = {}
        for tag, kw in (("full", {}), ("compute_only", {"copies": False}), ("copies_only", {"compute": False})):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            self._run_copies(FORWARD, slab, lvl, lvl, lvl, "I0", **kw)
            torch.cuda.synchronize()
            evs[0].record()
            for i in range(reps):
                self._run_copies(FORWARD, slab, lvl, lvl, lvl, "I%d" % (i & 1), **kw)
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
