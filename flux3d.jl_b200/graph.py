"""One submission per step: a sequence of library calls captured into a CUDA graph.

At BASELINE cfg4 sizes every mesh operator of the fit_mesh objective (examples/fit_mesh.jl:78-84: two sample_points, the chamfer
distance, laplacian_loss, edge_loss — forward and pullbacks) is one or two launches of 10-25 us that move less than a megabyte:
the step is bound by launch latency and by the host work between the launches, not by the kernels.  Every entry point of
libflux3d_b200 is asynchronous on the caller's stream and allocates nothing, so a whole step can be recorded once with stream
capture and replayed as ONE graph launch (CUDA.jl: ``CUDA.capture``; here: torch.cuda.CUDAGraph).  What has to vary between
replays lives in device memory: the vertices being optimised (a static tensor updated in place) and the sampling counter
(``sample_points(..., counter=...)``, f3d_sample_points_replayable)."""
import torch


class CapturedStep:
    """``step = capture_step(fn)``; ``out = step()`` replays.  ``fn`` takes no arguments, reads its inputs from tensors that stay
    alive (and are updated in place between replays), may call ``.backward()`` (gradients accumulate into ``.grad`` tensors
    that exist before the capture) and must not synchronise (no ``.item()``, no host reads)."""

    def __init__(self, fn, warmup: int = 3):
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):   # first calls set kernel attributes, build topologies, size workspaces
                fn()
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out


def capture_step(fn, warmup: int = 3) -> CapturedStep:
    return CapturedStep(fn, warmup)
