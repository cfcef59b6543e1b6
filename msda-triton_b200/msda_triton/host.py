"""Host-buffer entry point: MSDA forward (+ backward) on PINNED HOST tensors with the host<->device copies pipelined
against the kernels.

The reference has no equivalent (its tensors are expected on the device already); this is the end-to-end path
``bench.py`` reports as ``e2e``: the batch is cut into per-image chunks and three CUDA streams overlap

    H2D(chunk i+1)   ||   kernels(chunk i)   ||   D2H(chunk i-1)

so a step costs about max(H2D, D2H) + one chunk of compute instead of H2D + compute + D2H (PCIe is full duplex).
Chunks are whole images because ``grad_img`` couples all queries of one image.  Two sets of device staging and result buffers
are allocated once per :class:`HostMsda` and used alternately; a call allocates nothing.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels


def bind_host_thread_to_gpu(device: Optional[torch.device] = None) -> Optional[list]:
    """Restricts the calling process to the CPU cores NVML reports as local to ``device`` (same NUMA node / PCIe root),
    so that pinned host buffers allocated afterwards land in the memory next to that GPU.  With one process per GPU
    on a multi-socket host this is what keeps the per-GPU host<->device copies from all crossing the socket link.
    Returns the core list, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        props = torch.cuda.get_device_properties(dev)
        pynvml.nvmlInit()
        bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cores = [64 * i + b for i, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cores) & os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 -- an optimisation only
        return None


class HostMsda:
    def __init__(self, batch: int, num_pixels: int, heads: int, channels: int, queries: int, levels: int, points: int,
                 dtype: torch.dtype = torch.float32, device: Optional[torch.device] = None, backward: bool = True,
                 chunks: Optional[int] = None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = dict(dtype=dtype, device=self.device)

        def staging():
            return {
                "img": torch.empty((batch, num_pixels, heads, channels), **d),
                "pts": torch.empty((batch, queries, heads, levels, points, 2), **d),
                "aw": torch.empty((batch, queries, heads, levels, points), **d),
                "go": torch.empty((batch, queries, heads, channels), **d) if backward else None,
                # device-side results (copied out by the D2H stream; "drained" guards their reuse)
                "out": torch.empty((batch, queries, heads, channels), **d),
                "gimg": torch.empty((batch, num_pixels, heads, channels), **d) if backward else None,
                "gpts": torch.empty((batch, queries, heads, levels, points, 2), **d) if backward else None,
                "gaw": torch.empty((batch, queries, heads, levels, points), **d) if backward else None,
                "free": None,      # event: the kernels that read this staging set have finished
                "drained": None,   # event: the D2H copies out of this set's result buffers have finished
            }

        # two staging sets: the H2D copies of call i+1 never wait for the kernels of call i
        self._sets = [staging(), staging()]
        self._turn = 0
        self.batch = batch
        self.backward = backward
        # images per chunk: more chunks shorten the latency of ONE call (copy/compute overlap inside the call), fewer
        # chunks mean fewer, larger copies, which is what sustained back-to-back calls want (they overlap across calls)
        n_chunks = batch if chunks is None else max(1, min(int(chunks), batch))
        self._bounds = [(batch * c // n_chunks, batch * (c + 1) // n_chunks) for c in range(n_chunks)]
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)

    @staticmethod
    def _check_pinned(*tensors):
        for t in tensors:
            if t is not None and not (t.device.type == "cpu" and t.is_pinned() and t.is_contiguous()):
                raise ValueError("HostMsda expects contiguous pinned host tensors (torch.Tensor.pin_memory())")

    def run(self, img, img_shapes_dev, sampling_points, attention_weights, padding_mode, align_corners,
            out, out_grad=None, img_grad=None, sampling_points_grad=None, attention_weights_grad=None,
            deterministic: Optional[bool] = None):
        """All tensor arguments except ``img_shapes_dev`` (int64 [L,2] on the device) are pinned host tensors;
        ``out`` / ``*_grad`` are filled in place.  Returns after the last D2H copy has been ENQUEUED on an internal
        stream; call :meth:`synchronize` (or ``torch.cuda.synchronize()``) before reading the host buffers."""
        do_bwd = out_grad is not None
        if do_bwd and not self.backward:
            raise ValueError("this HostMsda was created with backward=False")
        self._check_pinned(img, sampling_points, attention_weights, out, out_grad, img_grad, sampling_points_grad,
                           attention_weights_grad)
        cur = torch.cuda.current_stream(self.device)
        st = self._sets[self._turn]
        self._turn ^= 1
        # A staging set is free as soon as the KERNELS of the call that used it last are done -- that call's D2H copies
        # may still be in flight, so back-to-back calls overlap this call's H2D with earlier D2H (PCIe is full duplex).
        if st["free"] is not None:
            self.h2d.wait_event(st["free"])
        if st["drained"] is not None:
            cur.wait_event(st["drained"])   # the kernels below overwrite this set's result buffers
        needs = (img_grad is not None, sampling_points_grad is not None, attention_weights_grad is not None)
        for lo, hi in self._bounds:
            sl = slice(lo, hi)
            with torch.cuda.stream(self.h2d):
                st["img"][sl].copy_(img[sl], non_blocking=True)
                st["pts"][sl].copy_(sampling_points[sl], non_blocking=True)
                st["aw"][sl].copy_(attention_weights[sl], non_blocking=True)
                if do_bwd:
                    st["go"][sl].copy_(out_grad[sl], non_blocking=True)
                ready = self.h2d.record_event()
            cur.wait_event(ready)
            o = kernels.b200_multi_scale_deformable_attention_fwd(
                st["img"][sl], img_shapes_dev, st["pts"][sl], st["aw"][sl], padding_mode, align_corners,
                out=st["out"][sl])
            grads = (None, None, None)
            if do_bwd and any(needs):
                grads = kernels.b200_multi_scale_deformable_attention_bwd(
                    st["go"][sl], st["img"][sl], img_shapes_dev, st["pts"][sl], st["aw"][sl], padding_mode, align_corners,
                    needs=needs, deterministic=deterministic, grads=(st["gimg"][sl], st["gpts"][sl], st["gaw"][sl]))
            done = cur.record_event()
            self.d2h.wait_event(done)
            with torch.cuda.stream(self.d2h):
                for dst, src in ((out, o), (img_grad, grads[0]), (sampling_points_grad, grads[1]),
                                 (attention_weights_grad, grads[2])):
                    if dst is not None and src is not None:
                        dst[sl].copy_(src, non_blocking=True)
        st["drained"] = self.d2h.record_event()
        st["free"] = cur.record_event()
        return out

    def synchronize(self) -> None:
        """Blocks the host until every result of the previous run() calls has landed in the host buffers."""
        self.d2h.synchronize()
