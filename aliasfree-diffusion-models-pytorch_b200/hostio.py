"""Host-resident tensors through the fused kernel: a three-stage pipeline
(H2D of chunk i+1 | kernel on chunk i | D2H of chunk i-1) on three CUDA streams with
double-buffered device staging, so that PCIe runs in both directions at once and the kernel hides
under the copies.  This is the call a user with data in (pinned) host memory makes; bench.py's
``e2e`` figure times it."""
import torch

from . import ops


class HostPipeline:
    def __init__(self, chunk_shape, dtype=torch.float32, device="cuda", slots=2):
        """``chunk_shape`` = [b, C, H, W] of one staged chunk (b images of the host batch)."""
        self.device = torch.device(device)
        self.slots, self.b = slots, chunk_shape[0]
        self.xin = [torch.empty(chunk_shape, dtype=dtype, device=self.device) for _ in range(slots)]
        self.yout = [torch.empty(chunk_shape, dtype=dtype, device=self.device) for _ in range(slots)]
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(device=self.device) for _ in range(3))

    def filtered_gelu(self, x_host, y_host, filt_up, filt_down):
        """``y_host[:] = filtered_gelu(x_host)`` chunk by chunk (both should be pinned).  Asynchronous
        with respect to the host: the caller's current stream waits for the last D2H copy."""
        ku, kd = ops.Taps(filt_up), ops.Taps(filt_down)
        B = x_host.shape[0]
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(cur)
        n = (B + self.b - 1) // self.b
        ev_k, ev_out = [None] * n, [None] * n
        for i in range(n):
            lo, hi = i * self.b, min(B, (i + 1) * self.b)
            slot = i % self.slots
            xin, yout = self.xin[slot][: hi - lo], self.yout[slot][: hi - lo]
            with torch.cuda.stream(self.s_in):
                if i >= self.slots:
                    self.s_in.wait_event(ev_k[i - self.slots])          # input slot has been consumed
                xin.copy_(x_host[lo:hi], non_blocking=True)
                ev_in = self.s_in.record_event()
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(ev_in)
                if i >= self.slots:
                    self.s_k.wait_event(ev_out[i - self.slots])         # output slot has been drained
                ops._fgelu_fwd(xin, None, ku, kd, out=yout)
                ev_k[i] = self.s_k.record_event()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_k[i])
                y_host[lo:hi].copy_(yout, non_blocking=True)
                ev_out[i] = self.s_out.record_event()
        cur.wait_stream(self.s_out)
        return y_host
