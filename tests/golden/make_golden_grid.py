"""Golden fixture for grid_weight_render, produced by the UNMODIFIED reference CUDA extension on a B200:
    gpurun -- python tests/golden/make_golden_grid.py gpurun_out/golden
Two cases on one 24^3 sigma grid: a world-space camera and a forward-facing NDC camera (rt_kernel.cu:1240-1344,
1454-1478). Outputs are deterministic (a float max and an integer-valued count per cell)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from svox_t_b200 import synth  # noqa: E402
import refdrv  # noqa: E402


def grid_case(ndc):
    """Shared with tests/test_gpu_parity.py::_grid_case (same seeds, same cameras)."""
    rng = np.random.default_rng(17)
    reso = 24
    grid = (rng.random((reso, reso, reso)) * 12.0 - 4.0).astype(np.float32)
    W, H, fx = 44, 31, 50.0
    if ndc:
        c2w = synth.look_at((0.1, -0.05, 3.0), target=(0.0, 0.0, 0.0))
        kw = dict(ndc_width=W, ndc_height=H, ndc_focal=fx)
        off, inv = np.full(3, 0.5, np.float32), np.full(3, 0.5, np.float32)
    else:
        c2w = synth.synth_cameras(1)[0]
        kw = {}
        off, inv = np.zeros(3, np.float32), np.ones(3, np.float32)
    return grid, np.asarray(c2w, np.float32), W, H, fx, off, inv, kw


if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_new")
    assert refdrv.available()
    dev = torch.device("cuda:0")
    m = refdrv.module()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = {}
    for tag, ndc in (("world", False), ("ndc", True)):
        grid, c2w, W, H, fx, off, inv, kw = grid_case(ndc)
        cam, opt = m.CameraSpec(), refdrv.options(sigma_thresh=0.5)
        cam.c2w, cam.fx, cam.fy, cam.width, cam.height = t(c2w), fx, fx, W, H
        for k, v in kw.items():
            setattr(opt, k, v)
        gw, gh = m.grid_weight_render(t(grid), cam, opt, t(off), t(inv))
        torch.cuda.synchronize()
        res[tag + "_weight"] = gw.cpu().numpy()
        res[tag + "_hit"] = gh.cpu().numpy().astype(np.uint16)
        assert gh.sum().item() > 1000
        print(tag, "hits", gh.sum().item(), "max weight", gw.max().item())
    res["grid"] = grid_case(False)[0]
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "y_grid_weight.npz"), **res)
