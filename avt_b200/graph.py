"""Whole-step CUDA graph: forward + loss + backward + gradient all-reduce + optimizer step captured once, replayed
per iteration. The AVT step is ~560 kernel launches (ViT-B/16 + AVT-h); enqueueing them from Python costs ~12 ms of
host time per step, which is on par with the device time and leaves the GPU idle inside the small-kernel stretches
(AVT-h, loss head). A replay costs one launch. The reference has no equivalent (eager PyTorch, func/train.py:204-236).

What makes the step capturable:
  * every avt_* kernel runs on torch's current stream and allocates nothing;
  * dropout masks come from Philox offsets kept in device memory (avt_epilogue_t.drop_offset_dev), advanced by a
    captured in-place add, so every replay draws fresh masks; torch's own nn.Dropout is graph-safe by itself;
  * weights, gradients, optimizer state and activations are static buffers (flat ParamPack / workspaces).
Inputs are copied into static device buffers before each replay (from pinned host memory in the end-to-end path).
"""
import torch

from . import _lib


class GraphedStep:
    def __init__(self, fn, example_inputs, warmup=3):
        """fn(*tensors) -> tensor (e.g. the loss). `example_inputs`: device tensors of the final shapes/dtypes; they
        become the static input buffers. `warmup` eager calls run first (lazy allocations, optimizer state)."""
        self.fn = fn
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.static_output = fn(*self.static_inputs)
        self.avt_launches = _lib.launch_count - l0   # avt_* kernels inside one replay

    def __call__(self, *inputs, after_copy=None):
        """Copy `inputs` into the static buffers, replay, return the static output. `after_copy()` runs between the input
        copies and the replay (e.g. to record the event that lets a side stream refill a staging buffer while the step
        computes: double-buffered input prefetch)."""
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        if after_copy is not None:
            after_copy()
        self.graph.replay()
        return self.static_output
