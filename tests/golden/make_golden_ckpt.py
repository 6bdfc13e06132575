"""Checkpoint fixture written by the REFERENCE's own Python (`svox_t.N3Tree.save`, svox.py:679-710), CPU only:
    cd /tmp && python /root/repo/tests/golden/make_golden_ckpt.py /root/repo/tests/golden/ref_ckpt_L3.npz
The tree is the reference's `N3Tree(init_reserve=...)` + 2 x refine() + a selective refine of the leaves around one
corner, with extra_data and a non-default cube. tests/test_host_cpu.py loads it with svox_t_b200 and compares it with
the same construction done through svox_t_b200's own refine()."""
import sys
import warnings

sys.path.insert(0, "/root/reference")
for k in list(sys.modules):
    if k.startswith("svox_t"):
        del sys.modules[k]
warnings.simplefilter("ignore")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import svox_t  # noqa: E402  (the reference package; its CUDA extension is absent here, which only disables queries)

assert svox_t.__file__.startswith("/root/reference"), svox_t.__file__
out = sys.argv[1]
extra = torch.arange(12, dtype=torch.float32).reshape(4, 3)
t = svox_t.N3Tree(N=2, data_dim=5, init_reserve=2000, depth_limit=6, radius=[0.8, 0.5, 0.4], center=[0.1, 0.2, 0.3],
                  data_format="SH1", extra_data=extra)
t.refine()
t.refine()          # (refine(repeats=2) raises in the reference: leaf_node is not re-derived for the second round)
# selective refine: the leaves of node 1 (first child of the root), given as (node, i, j, k) rows like tree[pts] does
leaf = torch.tensor([[1, 0, 0, 0], [1, 0, 0, 1], [1, 1, 1, 1]])
t.refine(sel=(*leaf.T,), leaf_node=leaf)
t.save(out, shrink=True, compress=True)
z = np.load(out)
print({k: (z[k].shape, z[k].dtype) for k in z.files}, "n_internal", int(z["n_internal"]))
