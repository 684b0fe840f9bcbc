"""CUDA-graph replay of a Caduceus forward (one process, one GPU).

A forward is ~25 launches per layer (this library's kernels through ctypes + cuBLAS GEMMs + a few tiny torch ops); at
short sequence lengths (<= 8k) the step is bound by launch/Python overhead, not by the GPU.  Every launch of this
package goes to `torch.cuda.current_stream()`, never allocates through its own allocator and never synchronises, so a
whole forward can be captured once and replayed from static buffers.
"""
import torch


class GraphedForward:
    """Capture `model(ids).logits` (no grad) for a fixed input shape and replay it.

        fwd = GraphedForward(model, example_ids)      # example_ids: (B, L) int64 CUDA tensor
        logits = fwd(ids)                             # same shape; returns the graph's static output tensor
    """

    def __init__(self, model, example_ids, warmup=3):
        if not example_ids.is_cuda:
            raise RuntimeError("GraphedForward needs CUDA tensors")
        self.static_ids = example_ids.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):          # fills the derived-weight / job-table caches, cuBLAS workspaces, ...
                model(self.static_ids)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_logits = model(self.static_ids).logits

    def __call__(self, ids):
        self.static_ids.copy_(ids, non_blocking=True)
        self.graph.replay()
        return self.static_logits
