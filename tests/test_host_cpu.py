"""CPU tests of the host logic: N3Tree structure ops, spec packing, options, persistence, and that the product
path FAILS LOUDLY (no CPU fallback, no oracle behind the API) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import svox_t_b200 as sv
from svox_t_b200 import synth


def test_refine_builds_full_tree_like_reference():
    t = sv.N3Tree(N=2, data_dim=16)
    for _ in range(3):
        t.refine()
    assert (t.filled, t.n_leaves, t.max_depth, t.capacity) == (585, 4096, 3, 585)
    ch, pd = t.child.reshape(-1, 8).numpy(), t.parent_depth.numpy()
    node, slot = np.nonzero(ch)
    kid = node + ch[node, slot]
    assert (pd[kid, 0] == node * 8 + slot).all() and (pd[kid, 1] == pd[node, 1] + 1).all()
    assert (t.data == sv.svox.EMPTY).all()


def test_init_refine_and_repeats_work():
    t = sv.N3Tree(N=2, data_dim=4, init_refine=2)           # the reference crashes here (Appendix B5)
    assert t.n_leaves == 512 and t.filled == 73
    t3 = sv.N3Tree(N=3, data_dim=4)
    t3.refine(repeats=2)
    assert t3.n_leaves == 27 ** 3 and t3.child.shape[1:] == (3, 3, 3)


def test_depth_limit_and_selected_refine():
    t = sv.N3Tree(N=2, data_dim=4, depth_limit=1)
    t.refine(); t.refine(); t.refine()
    assert t.max_depth == 1                                   # never deeper than depth_limit
    t = sv.N3Tree(N=2, data_dim=4)
    t.data[0, 1, 0, 1, 0] = 7
    leaf = torch.tensor([[0, 1, 0, 1]])
    t.refine(sel=(*leaf.T,), leaf_node=leaf)
    assert t.filled == 2 and int(t.child[0, 1, 0, 1]) == 1
    assert (t.data[1] == 7).all()                             # children inherit the parent's row (svox.py:539-540)
    assert t.parent_depth[1].tolist() == [5, 1]


def test_resize_keeps_sentinel_and_state_dict_names():
    t = sv.N3Tree(N=2, data_dim=4, init_reserve=1, geom_resize_fact=1.5)
    assert t.refine() is True
    assert (t.data[t.filled:] == sv.svox.EMPTY).all() and (t.child[t.filled:] == 0).all()
    names = set(t.state_dict().keys())
    assert {"features", "data", "child", "parent_depth", "_n_internal", "_n_free", "invradius", "offset"} <= names
    assert sv.svox.EMPTY == 1410065408


def test_world_transform_and_pack_index():
    t = sv.N3Tree(radius=[1.0, 2.0, 4.0], center=[0.0, 1.0, -1.0])
    p = torch.tensor([[0.0, 1.0, -1.0], [1.0, 3.0, 3.0]])
    assert torch.allclose(t.world2tree(p), torch.tensor([[0.5, 0.5, 0.5], [1.0, 1.0, 1.0]]))
    assert torch.allclose(t.tree2world(t.world2tree(p)), p)
    x = torch.tensor([[5, 1, 0, 1], [0, 0, 0, 0]])
    assert t._pack_index(x).tolist() == [45, 0]
    assert t._unpack_index(t._pack_index(x)).tolist() == x.tolist()


def test_from_tensors_and_save_load_roundtrip(tmp_path):
    tr = synth.synth_tree(3, "ball", r_out=0.4)
    t = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=8)
    assert t.filled == tr["n_nodes"] and t.n_leaves == tr["n_leaves"] and t.max_depth == 2
    path = str(tmp_path / "tree.npz")
    t.save(path)
    u = sv.N3Tree.load(path)
    for k in ("child", "data", "parent_depth", "invradius", "offset"):
        assert torch.equal(getattr(t, k), getattr(u, k))
    assert u.filled == t.filled and u.data_dim == 8 and repr(u.data_format) == "RGBA"


def test_data_format_parsing():
    assert (sv.DataFormat("RGBA").format, sv.DataFormat("RGBA").basis_dim) == (0, -1)
    assert (sv.DataFormat("SH9").format, sv.DataFormat("SH9").basis_dim) == (1, 9)
    assert repr(sv.DataFormat("SG25")) == "SG25" and repr(sv.DataFormat("ASG4")) == "ASG4"


def test_render_options_follow_reference_defaults():
    r = sv.VolumeRenderer(sv.N3Tree(data_dim=8))
    o = r._get_options()
    assert (o.step_size, o.background_brightness, o.sigma_thresh, o.stop_thresh, o.ndc_width) == (1e-3, 1.0, 0.0, 0.0, -1)
    o = r._get_options(fast=True)
    assert (o.sigma_thresh, o.stop_thresh) == (1e-2, 1e-2)
    r.sigma_thresh = 0.5
    assert r._get_options().sigma_thresh == 0.5               # instance override (renderer.py:435-438)


def test_no_cpu_fallback_anywhere():
    t = sv.N3Tree(data_dim=8, init_refine=1)
    r = sv.VolumeRenderer(t)
    rays = sv.Rays(torch.zeros(4, 3), torch.ones(4, 3), torch.ones(4, 3))
    with pytest.raises(RuntimeError):
        r(t.features, rays)
    with pytest.raises(RuntimeError):
        r(t.features, rays, cuda=False)
    with pytest.raises(RuntimeError):
        r.render_persp(t.features, torch.eye(4))
    with pytest.raises(RuntimeError):
        t(t.features, torch.rand(3, 3))
    with pytest.raises(RuntimeError):
        sv.csrc.volume_render(t._spec(t.features), sv.renderer._rays_spec_from_rays(rays), r._get_options())


def test_float64_specs_dispatch_and_refuse_cpu_tensors():
    """A double feature table selects the float64 instantiation (promoted offset / scaling, no accelerator or derived
    tables); like every other path it refuses CPU tensors instead of falling back."""
    t = sv.N3Tree(data_dim=8, init_refine=1)
    f64 = t.features.detach().double()
    ts = t._spec(f64)
    assert ts.is_f64 and ts.offset.dtype == torch.float64 and ts.scaling.dtype == torch.float64 and ts._accel is None
    assert not t._spec(t.features, _with_accel=False).is_f64
    assert torch.equal(ts.offset, t.offset.double())
    rays = sv.renderer._rays_spec_from_rays(sv.Rays(torch.zeros(4, 3).double(), torch.ones(4, 3).double(),
                                                    torch.ones(4, 3).double()))
    opt = sv.VolumeRenderer(t)._get_options()
    for call in (lambda: sv.csrc.volume_render(ts, rays, opt),
                 lambda: sv.csrc.render_depth(ts, rays, opt),
                 lambda: sv.csrc.volume_render_backward(ts, rays, opt, torch.zeros(4, 8).double()),
                 lambda: sv.csrc.query_vertical(ts, torch.rand(3, 3).double())):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()
    opt.format = sv.csrc.FORMAT_SH                               # view-dependent formats are float32 only
    with pytest.raises(RuntimeError, match="RGBA"):
        sv.csrc.volume_render(ts, rays, opt)
    assert ctypes.sizeof(sv.csrc._CTree64) == 72 and ctypes.sizeof(sv.csrc._CCamera64) == 40


def test_drop_in_alias_package():
    import svox_t
    import svox_t.csrc as _C
    assert svox_t.N3Tree is sv.N3Tree and svox_t.VolumeRenderer is sv.VolumeRenderer and svox_t.Rays is sv.Rays
    # the reference probes its extension exactly like this (svox_t/helpers.py:363-376)
    assert hasattr(_C, "query_vertical") and hasattr(_C, "volume_render") and hasattr(_C, "TreeSpec")


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "svox_t_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle\.oracle|svox_oracle|_ref/", re.M)
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not pat.search(src), f"{f} references the oracle"


def test_loads_a_checkpoint_written_by_the_reference(tmp_path):
    """tests/golden/ref_ckpt_L3.npz was written by the reference's own N3Tree.save (tests/golden/make_golden_ckpt.py).
    load() must read it, the same construction through this package's refine() must give identical tensors, and
    save() must write a file with the reference's keys / dtypes / shapes."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "ref_ckpt_L3.npz")
    z = np.load(path)
    t = sv.N3Tree.load(path)
    assert (t.N, t.data_dim, t.depth_limit, t.filled, t.capacity) == (2, 5, 6, 76, 76)
    assert repr(t.data_format) == "SH1" and torch.equal(t.extra_data, torch.arange(12.0).reshape(4, 3))
    assert torch.allclose(t.invradius, torch.tensor([0.5 / 0.8, 0.5 / 0.5, 0.5 / 0.4]))
    assert torch.allclose(t.offset, 0.5 * (1.0 - torch.tensor([0.1, 0.2, 0.3]) / torch.tensor([0.8, 0.5, 0.4])))
    mine = sv.N3Tree(N=2, data_dim=5, init_reserve=2000, depth_limit=6, radius=[0.8, 0.5, 0.4], center=[0.1, 0.2, 0.3],
                     data_format="SH1", extra_data=torch.arange(12.0).reshape(4, 3))
    mine.refine()
    mine.refine()
    leaf = torch.tensor([[1, 0, 0, 0], [1, 0, 0, 1], [1, 1, 1, 1]])
    mine.refine(sel=(*leaf.T,), leaf_node=leaf)
    n = mine.filled
    assert n == 76
    assert torch.equal(mine.child[:n], t.child) and torch.equal(mine.data[:n], t.data)
    assert torch.equal(mine.parent_depth[:n], t.parent_depth)
    out = tmp_path / "again.npz"
    mine.save(str(out))
    z2 = np.load(out)
    assert sorted(z2.files) == sorted(z.files)
    for k in z.files:
        assert z2[k].shape == z[k].shape and z2[k].dtype.kind == z[k].dtype.kind, k
        assert np.array_equal(z2[k], z[k]), k


def test_clone_shrink_and_shape_accessors():
    t = sv.N3Tree(N=2, data_dim=6, init_reserve=100, extra_data=torch.ones(2, 3))
    t.refine()
    t.refine()
    assert t.capacity == 100 and t.filled == 73 and (t.ndim, tuple(t.shape), t.size(0), t.size(1), t.numel()) == (2, (512, 6), 512, 6, 3072)
    u = t.clone()
    assert u is not t and u.filled == t.filled and torch.equal(u.child, t.child) and torch.equal(u.data, t.data)
    assert torch.equal(u.extra_data, t.extra_data) and u.features.data_ptr() != t.features.data_ptr()
    u.refine()
    assert u.filled > t.filled                                  # the copy is independent
    assert t.shrink_to_fit() and t.capacity == t.filled == 73 and not t.shrink_to_fit()
    with pytest.raises(RuntimeError):
        t.partial(data_sel=-1)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the CUDA arm) prints ONE JSON line with the
    contract's keys; it needs no GPU, so the contract is checked here."""
    import json
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(__file__), "..")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["value"] > 0 and j["unit"] == "Mrays/s" and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_view_row_accessors_on_cpu_tensors():
    """N3TreeView.values / set / in-place row ops are plain tensor indexing over `features` (no kernel involved), so they
    are checked here on a CPU tree assembled from the generator's tensors; anything that needs a kernel still raises."""
    import svox_t_b200 as sv
    tr = synth.synth_tree(3, "ball", r_out=0.45)
    D = 5
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location="cpu")
    f = torch.from_numpy(synth.synth_features(tr["M"], D))
    tree.features = torch.nn.Parameter(f.clone())
    view = tree[:]
    idx = tree.data[view.key][..., 0].long()
    ok = idx < tr["M"]
    assert 0 < int(ok.sum()) < len(view) and view.shape == (len(view), D) and view.ndim == 2
    vals = view.values
    assert torch.equal(vals[ok].detach(), f[idx[ok]]) and float(vals[~ok].detach().abs().sum()) == 0.0
    view.clamp_(min=0.25)
    assert float(tree.features.detach()[idx[ok]].min()) >= 0.25
    view.set(torch.full((len(view), D), 2.0))
    assert bool((tree.features.detach()[idx[ok]] == 2.0).all())
    view.uniform_(3.0, 4.0); view.relu_(); view.nan_to_num_()
    got = tree.features.detach()[idx[ok]]
    assert float(got.min()) >= 3.0 and float(got.max()) <= 4.0
    assert view._indexer().shape == (len(view), 4) and "leaves" in repr(view)
    with pytest.raises(RuntimeError):
        view.corners_local                       # calc_corners is a CUDA kernel: no CPU fallback
    with pytest.raises(RuntimeError):
        tree.set(torch.rand(4, 3), torch.rand(4, D))


def test_leaf_grad_exchange_single_rank_bookkeeping():
    """World of one (no process group): the exchange object is a plain table; the zero-fill bookkeeping hands the table
    out as it is exactly once after a table pass for THESE features, and zero-fills it otherwise."""
    import torch
    from svox_t_b200 import dist as svd
    x = svd.LeafGradExchange(6, 4, "cpu")
    assert x.world == 1 and x.backend == "local" and x.table.shape == (6, 4) and x.status() == 0
    f1, f2 = torch.randn(6, 4), torch.randn(6, 4)
    x.table.fill_(3.0)
    assert float(x.table_for_backward(f1).abs().sum()) == 0.0        # nothing was noted: zero-filled now
    x.table.fill_(5.0)
    x.note_zeroed(f1)                                                # "the table pass for f1 has just zeroed it"
    assert float(x.table_for_backward(f1).sum()) == 5.0 * 24          # handed out untouched ...
    assert float(x.table_for_backward(f1).abs().sum()) == 0.0        # ... once
    x.table.fill_(7.0)
    x.note_zeroed(f1)
    assert float(x.table_for_backward(f2).abs().sum()) == 0.0        # other features: not trusted
    x.table.fill_(1.0)
    x.note_zeroed(f1)
    f1.add_(1.0)                                                     # in-place update bumps the version: stale
    assert float(x.table_for_backward(f1).abs().sum()) == 0.0
    assert x.all_reduce_() is x.table and x.describe()["world"] == 1
