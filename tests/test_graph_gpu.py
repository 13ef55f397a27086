"""The fit_mesh objective (examples/fit_mesh.jl:78-84) — two sample_points, chamfer distance, laplacian_loss, edge_loss, forward and
pullbacks — captured once and replayed as ONE CUDA-graph launch: same numbers as the eager calls, fresh samples at every replay."""
import numpy as np
import pytest
import torch

from fixtures import teapots

pytestmark = pytest.mark.gpu


def _objective(f3d, src, tgt, delta, S, c1, c2):
    md = f3d.offset(src, delta)
    a = f3d.sample_points(md, S, seed=11, counter=c1)
    b = f3d.sample_points(tgt, S, seed=12, counter=c2)
    return f3d.chamfer_distance(a, b) + 0.1 * f3d.laplacian_loss(md) + f3d.edge_loss(md)


def test_fit_mesh_step_as_one_graph(f3d, oracle, golden_dir):
    vl, fl = teapots(4, golden_dir, oracle)
    src = f3d.TriMesh(vl, fl)
    tgt = f3d.TriMesh([v * np.float32(1.07) + np.float32(0.01) for v in vl], fl)
    S = 3000
    nV = sum(len(v) for v in vl)
    delta = torch.zeros((nV, 3), device="cuda", requires_grad=True)
    delta.grad = torch.zeros_like(delta)
    c1 = torch.zeros(1, dtype=torch.int64, device="cuda")
    c2 = torch.zeros(1, dtype=torch.int64, device="cuda")

    def step():
        delta.grad.zero_()
        loss = _objective(f3d, src, tgt, delta, S, c1, c2)
        loss.backward()
        return loss

    graphed = f3d.capture_step(step)
    counters_after_capture = (int(c1.item()), int(c2.item()))
    losses, grads = [], []
    for _ in range(3):
        out = graphed()
        torch.cuda.synchronize()
        losses.append(float(out.item()))
        grads.append(delta.grad.clone())
    # every replay advanced the device counters by one: fresh draws, hence (slightly) different losses
    assert (int(c1.item()), int(c2.item())) == (counters_after_capture[0] + 3, counters_after_capture[1] + 3)
    assert len(set(losses)) == 3
    # the eager step at the same counter state gives the same loss and (up to the RED.ADD order of the sampling pullback) gradient
    c1.fill_(counters_after_capture[0] + 2)
    c2.fill_(counters_after_capture[1] + 2)
    eager = step()
    torch.cuda.synchronize()
    assert abs(float(eager.item()) - losses[2]) <= 1e-6 * abs(losses[2])
    assert torch.allclose(delta.grad, grads[2], rtol=1e-4, atol=1e-7)
    assert float(delta.grad.abs().max()) > 0


def test_replayable_counter_changes_the_draws(f3d, oracle, golden_dir):
    vl, fl = teapots(2, golden_dir, oracle)
    m = f3d.TriMesh(vl, fl)
    c = torch.zeros(1, dtype=torch.int64, device="cuda")
    a = f3d.sample_points(m, 500, seed=5, counter=c)
    b = f3d.sample_points(m, 500, seed=5, counter=c)
    assert int(c.item()) == 2 and not torch.equal(a, b)
    assert torch.equal(a, f3d.sample_points(m, 500, seed=5, offset=0))
    assert torch.equal(b, f3d.sample_points(m, 500, seed=5, offset=1))
