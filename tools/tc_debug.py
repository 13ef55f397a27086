"""Development aid: where do the tensor-core sweep and the CUDA-core sweep disagree?  Prints every mismatching row with the
work item / CTA / candidate tile it belongs to and the locator triple the sweep left in the workspace."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d  # noqa: E402

B, N, M = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32x4096x4096").split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
A = torch.from_numpy(np.random.default_rng(201).random((B, N, 3), dtype=np.float32)).cuda()
Bc = torch.from_numpy(np.random.default_rng(202).random((B, M, 3), dtype=np.float32)).cuda()
ref = f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=f3d.FLAG_CUDA_CORES)
torch.cuda.synchronize()
refA, refB = ref[2].clone(), ref[3].clone()


def up(x, a):
    return (x + a - 1) // a * a


NpA, NpB = up(N, 256), up(M, 256)
rbA, rbB = NpA // 256, NpB // 256
for rep in range(reps):
    out = f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=f3d.FLAG_TENSOR)
    torch.cuda.synchronize()
    ws = f3d._lib.workspace(("chamfer", B, N, M), 256, A.device)
    hdr = ws[:12].cpu().numpy().view(np.int32)
    badA = (out[2] != refA).nonzero().cpu().numpy()
    badB = (out[3] != refB).nonzero().cpu().numpy()
    print(f"rep {rep}: loss {out[0].item():.9g} (ref {ref[0].item():.9g}); mismatches dir0 {len(badA)} dir1 {len(badB)}; amb {hdr[1]} viol {hdr[2]}")
    for d, bad, got, want in ((0, badA, out[2], refA), (1, badB, out[3], refB)):
        for b, q in bad[:40]:
            rb = q // 256
            item = b * (rbA + rbB) + (rbA if d else 0) + rb
            g, w = int(got[b, q]), int(want[b, q])
            print(f"   dir {d} b {b} row {q} (item {item} cta {item % 148} it {item // 148} rtile {(q % 256) // 128} lane {q % 128}): got {g} (tile {g // 128}) want {w} "
                  f"(tile {w // 128} chunk {w // 32})")
