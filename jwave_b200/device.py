"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out.

torch is plumbing here (device memory, streams, torch.distributed); the arithmetic is
libjwave_cuda.so's *_dev functions, enqueued on torch's current stream so torch.cuda.Event
brackets them.

One stream at a time per context: every call tells the library which torch stream is current
(jwc_set_stream); when that changes, the library makes the new stream wait on the device for the work
enqueued under the old one, because both use the context's scratch buffers (include/jwave_cuda.h).  Callers
that want two streams to overlap use two DeviceTransforms (two contexts)."""
import torch

from . import _lib
from .exceptions import JWaveFailure
from .transforms import CudaContext


class DeviceTransforms:
    """FWT / WPT over float64 CUDA tensors for one wavelet on one GPU."""

    def __init__(self, wavelet, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("jwave_b200.device needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.ctx = CudaContext(self.device.index)
        self.wid = self.ctx.register(wavelet)
        self.wavelet = wavelet
        self._L = self.ctx._lib

    def _prep(self, x, out):
        if x.dtype != torch.float64 or not x.is_cuda or x.device != self.device:
            raise JWaveFailure("expected a float64 tensor on " + str(self.device))
        x = x.contiguous()
        if out is None:
            out = torch.empty_like(x)
        elif out.shape != x.shape or out.dtype != x.dtype or out.device != x.device or not out.is_contiguous():
            raise JWaveFailure("out must match the input (contiguous float64, same device)")
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        return x, out

    def axis(self, kind, direction, x, outer, n, inner, level, out=None):
        """jwc_axis_dev on a dense [outer][n][inner] view of x."""
        x, out = self._prep(x, out)
        if outer * n * inner != x.numel():
            raise JWaveFailure("outer * n * inner must equal the number of elements")
        st = self._L.jwc_axis_dev(self.ctx.handle, self.wid, kind, direction, x.data_ptr(), out.data_ptr(),
                                  outer, n, inner, level)
        self.ctx.check(st, "jwc_axis_dev")
        return out

    def axis_remote(self, kind, direction, x, outer, n, inner, level, peers, mode, lg_seg, lg_hi=0,
                    outer_stride=0, row_stride=0, base_off=0):
        """jwc_axis_dev_remote: the pass's final output is stored straight into the peers' buffers
        (`peers`: device pointers valid on THIS GPU, one per rank) - see include/jwave_cuda.h."""
        if x.dtype != torch.float64 or not x.is_cuda or x.device != self.device or not x.is_contiguous():
            raise JWaveFailure("expected a contiguous float64 tensor on " + str(self.device))
        if outer * n * inner != x.numel():
            raise JWaveFailure("outer * n * inner must equal the number of elements")
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        m = _lib.RemoteMap()
        m.mode, m.world, m.lg_seg, m.lg_hi = mode, len(peers), lg_seg, lg_hi
        m.outer_stride, m.row_stride, m.base_off = outer_stride, row_stride, base_off
        for i, ptr in enumerate(peers):
            m.peer[i] = ptr
        st = self._L.jwc_axis_dev_remote(self.ctx.handle, self.wid, kind, direction, x.data_ptr(), outer, n, inner,
                                         level, m)
        self.ctx.check(st, "jwc_axis_dev_remote")

    def transform1d(self, kind, direction, x, level, out=None):
        """x: [batch][n]"""
        x, out = self._prep(x, out)
        n = x.shape[-1]
        fn = self._L.jwc_fwt1d_dev if kind == _lib.FWT else self._L.jwc_wpt1d_dev
        st = fn(self.ctx.handle, self.wid, direction, x.data_ptr(), out.data_ptr(), x.numel() // n, n, level)
        self.ctx.check(st, "jwc_1d_dev")
        return out

    def transform2d(self, kind, direction, x, lvlM, lvlN, out=None):
        """x: [batch][rows][cols] (or [rows][cols])"""
        x, out = self._prep(x, out)
        rows, cols = x.shape[-2:]
        fn = self._L.jwc_fwt2d_dev if kind == _lib.FWT else self._L.jwc_wpt2d_dev
        st = fn(self.ctx.handle, self.wid, direction, x.data_ptr(), out.data_ptr(), x.numel() // (rows * cols),
                rows, cols, lvlM, lvlN)
        self.ctx.check(st, "jwc_2d_dev")
        return out

    def transform3d(self, kind, direction, x, lvlP, lvlQ, lvlR, out=None):
        """x: [P][Q][R]"""
        x, out = self._prep(x, out)
        P, Q, R = x.shape
        fn = self._L.jwc_fwt3d_dev if kind == _lib.FWT else self._L.jwc_wpt3d_dev
        st = fn(self.ctx.handle, self.wid, direction, x.data_ptr(), out.data_ptr(), P, Q, R, lvlP, lvlQ, lvlR)
        self.ctx.check(st, "jwc_3d_dev")
        return out

    def forward_compress1d(self, kind, x, level, threshold, out=None):
        """jwc_forward1d_compress_dev: CompressorMagnitude(threshold).compress(forward(x)) for a [batch][n] tensor in one
        call (Compressor.java:97-110 behind FastWaveletTransform / WaveletPacketTransform.forward); returns
        (coefficients with the small ones zeroed, magnitude as a 1-element device tensor)."""
        x, out = self._prep(x, out)
        n = x.shape[-1]
        mag = torch.empty(1, dtype=torch.float64, device=self.device)
        st = self._L.jwc_forward1d_compress_dev(self.ctx.handle, self.wid, kind, x.data_ptr(), out.data_ptr(),
                                                x.numel() // n, n, level, float(threshold), mag.data_ptr())
        self.ctx.check(st, "jwc_forward1d_compress_dev")
        return out, mag

    def copy2d(self, dst, dpitch, src, spitch, width, height, stream=None):
        """jwc_copy2d_dev: strided device copy on the copy engines; dst / src are device pointers (ints), pitches
        and width in bytes; `stream` a torch stream (default: the current one)."""
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        rc = self._L.jwc_copy2d_dev(self.ctx.handle, dst, dpitch, src, spitch, width, height, st)
        self.ctx.check(rc, "jwc_copy2d_dev")

    def launch_count(self):
        return self.ctx.launch_count()

    def close(self):
        self.ctx.close()
