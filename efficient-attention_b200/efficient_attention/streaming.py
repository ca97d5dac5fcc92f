"""Host-resident batches: overlap H2D copy, the module's forward and D2H copy chunk by chunk.

The attention core is far faster than PCIe, so a serving / evaluation loop that keeps activations in pinned
host memory is bound by the copies; splitting the batch and running the three stages on three CUDA streams
keeps both PCIe directions busy at once (full duplex) and hides the kernels under them.
"""
import torch


class HostPipeline:
    """`pipe(x_host, y_host)` computes `y_host[...] = module(x_host)` for pinned host tensors.

    module: any nn.Module on a CUDA device whose forward maps [B, ...] -> [B, ...] batch-wise.
    chunk:  images per stage; depth: device-side buffers per stage (2 = double buffering).
    """

    def __init__(self, module, chunk, depth=2):
        self.module = module
        self.chunk = int(chunk)
        self.depth = int(depth)
        p = next(module.parameters())
        self.device = p.device
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._x = [None] * self.depth
        self._ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        self._ev_comp = [torch.cuda.Event() for _ in range(self.depth)]
        self._ev_x_free = [None] * self.depth
        self._ev_out = [None] * self.depth

    @torch.no_grad()
    def __call__(self, x_host, y_host):
        assert x_host.is_pinned() and y_host.is_pinned(), 'HostPipeline needs pinned host tensors'
        B = x_host.shape[0]
        main = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(main)
        n_chunks = (B + self.chunk - 1) // self.chunk
        for c in range(n_chunks):
            lo, hi = c * self.chunk, min(B, (c + 1) * self.chunk)
            s = c % self.depth
            if self._x[s] is None or self._x[s].shape[0] < hi - lo or self._x[s].dtype != x_host.dtype:
                self._x[s] = torch.empty((self.chunk,) + tuple(x_host.shape[1:]), dtype=x_host.dtype, device=self.device)
            xd = self._x[s][:hi - lo]
            with torch.cuda.stream(self.s_in):
                if self._ev_x_free[s] is not None:
                    self.s_in.wait_event(self._ev_x_free[s])      # forward of chunk c-depth has consumed this buffer
                xd.copy_(x_host[lo:hi], non_blocking=True)
                self._ev_in[s].record(self.s_in)
            main.wait_event(self._ev_in[s])
            y = self.module(xd)
            self._ev_comp[s].record(main)
            self._ev_x_free[s] = self._ev_comp[s]
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self._ev_comp[s])
                y.record_stream(self.s_out)
                y_host[lo:hi].copy_(y, non_blocking=True)
        main.wait_stream(self.s_out)
        return y_host
