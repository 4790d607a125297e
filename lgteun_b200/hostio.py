"""End-to-end execution with HOST tensors: the eval loop around the hot path.

The reference's test loop moves a batch to the GPU, runs the forward and copies the result back
(`set_batch_cuda` -> `get_model_output` -> `torch2np`, models/base/base_model.py:293-305), strictly one after the
other.  Here the batch is cut into chunks and the three phases run on three CUDA streams, so that the PCIe copies of
chunk i+1 / i-1 hide under the kernels of chunk i (H2D and D2H are full duplex).  Only the upload of the first chunk and
the download of the last one are exposed, so those two chunks are a quarter of the regular size (`plan_chunks`).
Results are identical to `net(ms.cuda(), pan.cuda()).cpu()` — image pairs are independent."""
from __future__ import annotations

import torch


def plan_chunks(n: int, chunk: int) -> list:
    """[(lo, hi)] covering range(n): regular chunks of `chunk` pairs between a short first and a short last chunk
    (chunk // 4) whose copies cannot hide under another chunk's kernels.  Batches of at most one chunk are not cut."""
    chunk = max(int(chunk), 1)
    if n <= chunk:
        return [(0, n)] if n > 0 else []
    edge = max(chunk // 4, 1)
    cuts = [0, edge]
    while n - cuts[-1] > chunk + edge:
        cuts.append(cuts[-1] + chunk)
    if n - cuts[-1] > edge:
        cuts.append(n - edge)
    cuts.append(n)
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


class HostPipeline:
    def __init__(self, net, device=None, chunk: int = 64):
        self.net = net
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the module on a CUDA device")
        self.chunk = int(chunk)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._bufs = None

    def _buffers(self, ms, pan):
        key = (ms.shape[1:], pan.shape[1:], self.chunk)
        if self._bufs is None or self._bufs[0] != key:
            mk = lambda shp: [torch.empty((self.chunk, *shp), device=self.device) for _ in range(2)]
            self._bufs = (key, mk(ms.shape[1:]), mk(pan.shape[1:]))
        return self._bufs[1], self._bufs[2]

    @torch.no_grad()
    def __call__(self, ms_host: torch.Tensor, pan_host: torch.Tensor, out_host: torch.Tensor | None = None) -> torch.Tensor:
        """ms_host [N,B,h,w], pan_host [N,1,4h,4w] on the CPU (pinned for asynchronous copies); returns out_host
        [N,B,4h,4w] on the CPU (pinned if allocated here).  Synchronises before returning."""
        n = ms_host.shape[0]
        if out_host is None:
            out_host = torch.empty((n, ms_host.shape[1], 4 * ms_host.shape[2], 4 * ms_host.shape[3]), pin_memory=True)
        ms_d, pan_d = self._buffers(ms_host, pan_host)
        compute = torch.cuda.current_stream(self.device)
        free = [None, None]                 # event: staging buffer i may be overwritten (its forward has been enqueued and run)
        self.s_in.wait_stream(compute)
        k = 0
        for lo, hi in plan_chunks(n, self.chunk):
            b = k & 1
            with torch.cuda.stream(self.s_in):
                if free[b] is not None:
                    self.s_in.wait_event(free[b])
                ms_d[b][: hi - lo].copy_(ms_host[lo:hi], non_blocking=True)
                pan_d[b][: hi - lo].copy_(pan_host[lo:hi], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.s_in)
            compute.wait_event(ready)
            out = self.net(ms_d[b][: hi - lo], pan_d[b][: hi - lo])
            done = torch.cuda.Event()
            done.record(compute)
            free[b] = done
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                out_host[lo:hi].copy_(out, non_blocking=True)
                out.record_stream(self.s_out)
            k += 1
        compute.wait_stream(self.s_out)
        torch.cuda.synchronize(self.device)
        return out_host
