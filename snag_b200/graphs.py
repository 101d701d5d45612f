"""CUDA-graph capture of a whole forward + backward step.

At the reference's own batch sizes (B = 3500, 10-14 icl_loss calls per step) the loss-layer slice is launch bound:
~3 ms of kernels inside ~15 ms of Python / autograd / launch overhead. Every entry point of libsnag_b200.so only
enqueues work on the stream it is given (no allocation, no synchronisation, TMA descriptors passed by value), so a
step built on them can be captured once and replayed with a single graph launch.

    step = GraphedStep(lambda: layer(streams, hidden, joint, joint_fz, links, weight_norm), leaves)
    loss = step()            # replays forward + backward; gradients are in leaf.grad (static buffers)

The callable must read its inputs from fixed tensors (update them in place between replays, e.g.
`links.copy_(new_batch)`), must not synchronise with the host (no .item(), no numpy inputs) and must have a fixed
control flow. Same contract as torch.cuda.graphs' whole-network capture.
"""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, fn, leaves, warmup: int = 3):
        """fn() -> scalar loss tensor built from `leaves` (tensors with requires_grad=True)."""
        self.leaves = [t for t in leaves if t is not None]
        dev = self.leaves[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):              # lazy initialisations (cuBLAS handles, kernel attributes) happen here
                for t in self.leaves:
                    t.grad = None
                fn().backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        for t in self.leaves:
            t.grad = None                                  # gradients are allocated inside the graph's private pool
        with torch.cuda.graph(self.graph):
            self.loss = fn()
            self.loss.backward()
        self.grads = [t.grad for t in self.leaves]

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss
