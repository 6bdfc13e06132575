"""N3Tree -- the sparse voxel N^3-tree container behind the octree volume-rendering hot path.

Public surface and tensor format follow the reference (svox_t/svox.py): ``data`` is an int32 table of rows into
an external ``features[M, D]`` tensor (empty leaf = int(1e10) wrapped to int32, svox.py:123-124), ``child`` holds
relative child offsets (0 = leaf), ``parent_depth`` = (packed parent slot, depth). Buffer / parameter names match
the reference's ``state_dict``, so checkpoints interchange.

Only what the hot path needs is here: construction, ``refine`` (svox.py:488-560), point query
(svox.py:216-285 -> query_vertical), ``construct_tree`` (svox.py:160-161), ``_spec`` (svox.py:899-925), npz
save/load (svox.py:679-752) and the LBS helpers (svox.py:971-981). The legacy svox accessors that treat ``data``
as floats (set/snap/merge/partial/...) are out of scope.

New, B200-side: the tree caches a packed grid+brick accelerator (``tree.accel()``) that the march kernels walk;
it is rebuilt whenever ``child``/``data`` change (tensor version counters + shapes are the cache key).
"""
from warnings import warn

import numpy as np
import torch
from torch import autograd, nn

from . import csrc as _C
from .helpers import DataFormat, LocalIndex, N3TreeView

EMPTY = int(np.array(int(1e10)).astype(np.int32))  # 1410065408, the reference's wrapped sentinel (svox.py:124)


class _QueryVerticalFunction(autograd.Function):
    """svox.py:38-56. The backward is the row scatter-add the reference's query_vertical_backward states (its own
    kernel faults, Appendix B1)."""

    @staticmethod
    def forward(ctx, data, tree_spec, indices):
        out, node_ids, data_ids, leaf_node = _C.query_vertical(tree_spec, indices)
        ctx.mark_non_differentiable(node_ids, data_ids, leaf_node)
        ctx.tree_spec = tree_spec
        ctx.save_for_backward(indices)
        return out, node_ids, data_ids, leaf_node

    @staticmethod
    def backward(ctx, grad_out, *_unused):
        if not ctx.needs_input_grad[0]:
            return None, None, None
        return _C.query_vertical_backward(ctx.tree_spec, ctx.saved_tensors[0], grad_out.contiguous()), None, None


class _WarpVerticalFunction(autograd.Function):
    """svox.py:58-75."""

    @staticmethod
    def forward(ctx, transformation_matrix, coordinates, skinning_weights, joint_index):
        vertices, matrices = _C.warp_vertices(transformation_matrix, coordinates, skinning_weights, joint_index)
        ctx.save_for_backward(transformation_matrix, coordinates, skinning_weights, joint_index)
        return vertices, matrices

    @staticmethod
    def backward(ctx, vertices_grad_out, matrices_grad_out):
        T, x, w, ji = ctx.saved_tensors
        if vertices_grad_out is None:
            vertices_grad_out = torch.zeros_like(x)
        if matrices_grad_out is None:
            matrices_grad_out = torch.zeros(x.shape[0], 4, 4, dtype=x.dtype, device=x.device)
        g_x, g_T, g_w = _C.warp_vertices_backward(T, x, w, ji, vertices_grad_out.contiguous(),
                                                  matrices_grad_out.contiguous())
        return g_T, g_x, g_w, None


class N3Tree(nn.Module):
    """PyTorch N^3-tree (N=2: octree) with B200 CUDA kernels behind it. See module docstring."""

    def __init__(self, N=2, data_dim=4, depth_limit=10, init_reserve=1, init_refine=0, geom_resize_fact=1.5,
                 radius=0.5, center=[0.5, 0.5, 0.5], data_format="RGBA", extra_data=None, map_location="cpu"):
        super().__init__()
        assert N >= 2
        assert depth_limit >= 0
        self.N = int(N)
        self.data_dim = int(data_dim)
        if init_refine > 0:
            for i in range(1, init_refine + 1):
                init_reserve += (N ** i) ** 3
        dev = map_location
        self.register_parameter("features", nn.Parameter(torch.zeros(init_reserve, data_dim, device=dev)))
        self.register_buffer("data", torch.full((init_reserve, N, N, N, 1), EMPTY, dtype=torch.int32, device=dev))
        self.register_buffer("child", torch.zeros(init_reserve, N, N, N, dtype=torch.int32, device=dev))
        self.register_buffer("parent_depth", torch.zeros(init_reserve, 2, dtype=torch.int32, device=dev))
        self.register_buffer("_n_internal", torch.tensor(1, device=dev))
        self.register_buffer("_n_free", torch.tensor(0, device=dev))
        if isinstance(radius, (float, int)):
            radius = [radius] * 3
        radius = torch.tensor(radius, dtype=torch.float32, device=dev)
        center = torch.tensor(center, dtype=torch.float32, device=dev)
        self.register_buffer("invradius", 0.5 / radius)
        self.register_buffer("offset", 0.5 * (1.0 - center / radius))
        self.depth_limit = depth_limit
        self.geom_resize_fact = geom_resize_fact
        self.data_format = DataFormat(data_format) if data_format is not None else None
        if extra_data is not None:
            assert isinstance(extra_data, torch.Tensor)
            self.register_buffer("extra_data", extra_data.to(device=dev))
        else:
            self.extra_data = None
        self._ver = 0
        self._invalidate()
        self._lock_tree_structure = False
        self._weight_accum = None
        self._accel_cache = None
        self.filled = 1          # python mirror of _n_internal: no device sync on the hot path
        self.refine(repeats=init_refine)

    # ---- structure -------------------------------------------------------------------------------------------
    @classmethod
    def from_tensors(cls, child, data, parent_depth, data_dim, n_internal=None, radius=0.5,
                     center=[0.5, 0.5, 0.5], depth_limit=10, data_format="RGBA", map_location="cpu"):
        """Adopt tensors already in the reference format (what the reference's load() does, svox.py:726-739)."""
        child = torch.as_tensor(child)
        N = child.shape[1]
        tree = cls(N=N, data_dim=data_dim, depth_limit=depth_limit, radius=radius, center=center,
                   data_format=data_format, map_location=map_location)
        dev = tree.child.device
        tree.child = child.to(device=dev, dtype=torch.int32).contiguous()
        tree.data = torch.as_tensor(data).to(device=dev, dtype=torch.int32).reshape(*tree.child.shape, 1).contiguous()
        tree.parent_depth = torch.as_tensor(parent_depth).to(device=dev, dtype=torch.int32).contiguous()
        n = int(tree.child.shape[0] if n_internal is None else n_internal)
        tree._n_internal.fill_(n)
        tree.filled = n
        tree._invalidate()
        return tree

    def build_from_points(self, points, depth, capacity=None):
        """Rebuild the whole tree in one shot: finest level ``depth``, one depth-``depth`` leaf per occupied cell,
        leaf row = index of the (last) point inside it. Equivalent to (depth-1) x ``tree[points].refine()`` followed by
        ``construct_tree(points)`` (the per-frame rebuild of svox.py:160-161,488-560), isomorphic result.

        ``capacity`` (nodes): build into tensors of that many nodes WITHOUT any host synchronisation -- the per-frame
        path. ``tree.filled`` then stays unknown until first asked for (one read-back, raising if the capacity was too
        small); the march, the accelerator and the point query only need the tensors."""
        assert self.N == 2, "the one-shot builder is octree-only"
        if self._lock_tree_structure:
            raise RuntimeError("Tree locked")
        if depth - 1 > self.depth_limit:
            raise RuntimeError("depth exceeds depth_limit")
        child, data, parent_depth, status = _C.build_octree(points, depth, self.offset, self.invradius, capacity=capacity)
        self.child, self.data, self.parent_depth = child, data, parent_depth
        if status is None:
            self.filled = int(child.shape[0])
            self._n_internal.fill_(self.filled)
        else:
            self._filled_pending = status            # resolved by the `filled` property on first use
            self._n_internal.copy_(status[0].clamp(max=child.shape[0]))
        self._invalidate()
        self._known_depth = int(depth)     # spares the accelerator build its max-depth reduction + read-back
        return self

    @property
    def filled(self):
        """Number of nodes in use. After a capacity-bounded rebuild the count lives on the device until asked for."""
        st = self.__dict__.get("_filled_pending")
        if st is not None:
            need, over = (int(v) for v in st.tolist())
            self.__dict__["_filled_pending"] = None
            if over:
                raise RuntimeError(f"build_from_points: capacity {self.child.shape[0]} is too small, {need} nodes needed")
            self.__dict__["_filled"] = need
        return self.__dict__.get("_filled", 0)

    @filled.setter
    def filled(self, v):
        self.__dict__["_filled_pending"] = None
        self.__dict__["_filled"] = int(v)

    def _filled_bound(self):
        """Nodes in use, or -- while the exact count is still on the device -- the capacity (an upper bound: rows beyond
        the nodes in use are initialised, unreachable empty nodes)."""
        if self.__dict__.get("_filled_pending") is not None:
            return int(self.child.shape[0])
        return self.filled

    def construct_tree(self, indices):
        """data[leaf(p_i)] = i: point i becomes the feature row of its leaf (svox.py:160-161)."""
        _C.construct_tree(self._spec(self.features, _with_accel=False), indices)   # walks child/data (descend_ref)
        self._invalidate()

    def set(self, indices, values, cuda=True):
        """features[row of leaf(p_q), :K] = values[q] (svox.py:164-214 -> assign_vertical). Where several points share
        a leaf the last one (largest q) is taken. No autograd through ``indices`` or ``values``."""
        assert len(indices.shape) == 2
        assert not indices.requires_grad and not values.requires_grad
        if not cuda or not self.data.is_cuda:
            raise RuntimeError("svox_t_b200 has no CPU assignment path: the tree must be on a CUDA device")
        indices = indices.to(device=self.data.device, dtype=torch.float32).contiguous()
        values = values.to(device=self.data.device, dtype=torch.float32).contiguous()
        _C.assign_vertical(self._spec(self.features, _with_accel=False), indices, values)

    def _calc_corners(self, nodes, cuda=True):
        """Lower corners (tree coordinates) of the cells ``nodes[Q, 4] = [node, i, j, k]`` (svox.py:804-826)."""
        if not cuda or not self.data.is_cuda:
            raise RuntimeError("svox_t_b200 has no CPU path: the tree must be on a CUDA device")
        return _C.calc_corners(self._spec(self.features, _with_accel=False),
                               nodes.to(device=self.data.device, dtype=torch.int64).contiguous())

    def forward(self, features, indices, cuda=True, want_node_ids=False, world=True, want_data_ids=False,
                want_leaf_node=False):
        """Query leaf rows at points (Q, 3). Differentiable w.r.t. ``features`` (svox.py:216-285)."""
        assert not indices.requires_grad
        assert len(indices.shape) == 2
        if not cuda or not self.data.is_cuda:
            raise RuntimeError("svox_t_b200 has no CPU query path: the tree and the points must be on a CUDA device")
        result, node_ids, data_ids, leaf_node = _QueryVerticalFunction.apply(
            features, self._spec(features, world=world, _with_accel=False), indices)   # the point query walks child/data
        ret = [result, node_ids] if want_node_ids else result
        if want_data_ids:
            ret = ret if isinstance(ret, list) else [ret]
            ret.append(data_ids)
        if want_leaf_node:
            ret = ret if isinstance(ret, list) else [ret]
            ret.append(leaf_node)
        return ret

    def refine(self, repeats=1, sel=None, leaf_node=None, node_id=None):
        """Split the selected leaves (all leaves below depth_limit by default); svox.py:488-560.

        ``sel``: tuple of 4 index tensors (node, i, j, k) of unique leaves, ``leaf_node``: the same as [n,4].
        Returns True iff capacity grew. Unlike the reference (Appendix B5) ``repeats > 1`` works: without a selection
        every pass splits every leaf below depth_limit; with one, later passes split the children the previous pass
        created (all N^3 slots of the new nodes).
        """
        if self._lock_tree_structure:
            raise RuntimeError("Tree locked")
        resized = False
        explicit = sel is not None or leaf_node is not None
        with torch.no_grad():
            for repeat_id in range(repeats):
                filled = self.filled
                if sel is None and leaf_node is not None:
                    sel = (*leaf_node.T,)
                if sel is None:
                    leaves = self._all_leaves().to(self.data.device)
                    depths = self.parent_depth[leaves[:, 0], 1]
                    leaf_node = leaves[depths < self.depth_limit]
                    sel = (*leaf_node.T,)
                elif leaf_node is None:
                    leaf_node = torch.stack(sel, dim=-1).to(device=self.data.device)
                leaf_node = leaf_node.to(device=self.data.device, dtype=torch.int64)
                sel = tuple(t.to(device=self.data.device, dtype=torch.int64) for t in sel)
                num_nc = int(leaf_node.shape[0])
                if num_nc == 0:
                    return resized
                new_filled = filled + num_nc
                cap_needed = new_filled - self.capacity
                if cap_needed > 0:
                    self._resize_add_cap(cap_needed)
                    resized = True
                new_idxs = torch.arange(filled, new_filled, device=self.data.device, dtype=torch.int32)
                self.child[sel] = new_idxs - leaf_node[:, 0].to(torch.int32)
                self.data[filled:new_filled] = self.data[sel][:, None, None, None]        # children inherit the row
                self.parent_depth[filled:new_filled, 0] = (
                    self._pack_index(leaf_node) if node_id is None else node_id).to(torch.int32)
                self.parent_depth[filled:new_filled, 1] = self.parent_depth[leaf_node[:, 0], 1] + 1
                self._n_internal += num_nc
                self.filled += num_nc
                self._invalidate()
                if repeat_id + 1 < repeats:
                    if explicit:
                        # further repeats split the children just created (nodes [filled, new_filled)), as the
                        # reference infers its selector (svox.py:540-550), not every leaf of the tree
                        N = self.N
                        kids = torch.arange(filled, new_filled, device=self.data.device, dtype=torch.int64)
                        kids = kids[self.parent_depth[kids, 1] < self.depth_limit]
                        grid = torch.stack(torch.meshgrid(*(torch.arange(N, device=self.data.device),) * 3,
                                                          indexing="ij"), dim=-1).reshape(-1, 3)
                        leaf_node = torch.cat((kids.repeat_interleave(N ** 3)[:, None], grid.repeat(kids.shape[0], 1)), dim=1)
                        sel = (*leaf_node.T,)
                    else:
                        sel = leaf_node = None          # whole-tree refinement: every leaf again
                    node_id = None
        return resized

    # ---- bookkeeping -----------------------------------------------------------------------------------------
    @property
    def n_leaves(self):
        return self._all_leaves().shape[0]

    @property
    def n_internal(self):
        return self.filled

    @property
    def capacity(self):
        return self.parent_depth.shape[0]

    @property
    def max_depth(self):
        """Maximum tree depth - 1 (the reference's convention, svox.py:657-662)."""
        return int(self.parent_depth[:self.filled, 1].max().item())

    @property
    def depths(self):
        return self[:].depths

    def _pack_index(self, txyz):
        N = self.N
        return txyz[:, 0] * (N ** 3) + txyz[:, 1] * (N ** 2) + txyz[:, 2] * N + txyz[:, 3]

    def _unpack_index(self, flat):
        t = []
        for _ in range(3):
            t.append(flat % self.N)
            flat = flat // self.N
        return torch.stack((flat, t[2], t[1], t[0]), dim=-1)

    def _resize_add_cap(self, cap_needed):
        """Grow child/data/parent_depth geometrically (svox.py:841-863) -- on the device, without the reference's
        CPU bounce and torch.cuda.synchronize()."""
        cap_needed = max(cap_needed, int(self.capacity * (self.geom_resize_fact - 1.0)))
        dev = self.data.device
        self.data = torch.cat((self.data, torch.full((cap_needed, *self.data.shape[1:]), EMPTY,
                                                     dtype=self.data.dtype, device=dev)), dim=0)
        self.child = torch.cat((self.child, torch.zeros((cap_needed, *self.child.shape[1:]),
                                                        dtype=self.child.dtype, device=dev)))
        self.parent_depth = torch.cat((self.parent_depth, torch.zeros((cap_needed, 2),
                                                                      dtype=self.parent_depth.dtype, device=dev)))

    def _all_leaves(self):
        if self._last_all_leaves is None:
            self._last_all_leaves = (self.child[:self.filled] == 0).nonzero(as_tuple=False)
        return self._last_all_leaves

    def world2tree(self, indices):
        return torch.addcmul(self.offset, indices, self.invradius)

    def tree2world(self, indices):
        return (indices - self.offset) / self.invradius

    def _invalidate(self):
        self._known_depth = 0
        self._ver += 1
        self._last_all_leaves = None
        # a stale accelerator is kept as a shell: the per-frame rebuild refills its allocations in place
        self._accel_stale, self._accel_cache = getattr(self, "_accel_cache", None) or getattr(self, "_accel_stale", None), None

    # ---- the bridge to the kernels ---------------------------------------------------------------------------
    def accel(self, features=None, max_depth=0):
        """Packed grid+brick accelerator for the current child/data (N == 2 only; None otherwise). Cached."""
        if self.N != 2 or not self.data.is_cuda:
            return None
        feats = self.features if features is None else features
        spec = self._spec(feats, _with_accel=False)
        acc = self._accel_cache
        if acc is None or not acc.matches(spec):
            depth = max_depth or getattr(self, "_known_depth", 0)
            stale, self._accel_stale = getattr(self, "_accel_stale", None), None
            if stale is not None and depth and stale.rebuild(spec, depth):      # same allocations, no host sync
                self._accel_cache = stale
                return stale
            del stale
            try:
                acc = _C.Accel(spec, max_depth=depth)
            except RuntimeError as e:      # e.g. more rows than the packed index field can hold
                warn(f"svox_t_b200: accelerator not built ({e}); walking the reference tensors instead")
                acc = None
            self._accel_cache = acc
        return acc

    def activated(self, features, accel=None, grad_exchange=None):
        """Table of ``features`` with the sigmoid applied once per row (cached until ``features`` changes). The pass that
        builds it also refreshes ``accel``'s hit marks (csrc.Activated); ``grad_exchange``'s gradient table is zero-filled
        on a side stream for the backward of the step that starts with these features (it overlaps the forward)."""
        act = getattr(self, "_act_cache", None)
        if act is None or not act.matches(features):
            act = _C.Activated(features, accel=accel)
            if grad_exchange is not None:          # the gradient table of this step's backward: zeroed beside the forward
                grad_exchange.zero_async(features)
            self._act_cache = act
        elif accel is not None:
            accel.mark_hits(features)
        return act

    def sigma_table(self, features):
        """Compact sigma array of ``features`` (cached until ``features`` changes) for the sigma-only marches."""
        st = getattr(self, "_sigma_cache", None)
        if st is None or not st.matches(features):
            st = self._sigma_cache = _C.SigmaTable(features)
        return st

    def _spec(self, features, joint_features=None, skinning_weights=None, joint_index=None,
              transformation_matrices=None, world=True, _with_accel=True):
        """Pack the tree into a TreeSpec (svox.py:899-925). transformation_matrices / joint_* are carried for
        signature parity; the feature-level (RGBA) path ignores them, as the reference does."""
        dev = self.data.device
        ts = _C.TreeSpec()
        ts.features = features
        ts.data = self.data
        ts.child = self.child
        ts.parent_depth = self.parent_depth
        ts.extra_data = self.extra_data if self.extra_data is not None else torch.empty((0, 0), device=dev)
        if world:
            ts.offset, ts.scaling = self.offset, self.invradius
        else:
            if getattr(self, "_unit_xform", None) is None or self._unit_xform[0].device != dev:
                self._unit_xform = (torch.zeros(3, device=dev), torch.ones(3, device=dev))
            ts.offset, ts.scaling = self._unit_xform
        ts.n_internal = self._filled_bound()
        ts._weight_accum = self._weight_accum
        ts.joint_features, ts.skinning_weights, ts.joint_index = joint_features, skinning_weights, joint_index
        ts.transformation_matrices = transformation_matrices
        if features is not None and features.dtype == torch.float64:
            # float64 instantiation (the reference's AT_DISPATCH_FLOATING_TYPES): the general kernels walk child / data
            # in double; offset / scaling are promoted, no accelerator or derived table is attached
            ts.offset, ts.scaling = ts.offset.double(), ts.scaling.double()
            return ts
        if _with_accel:
            ts._accel = self.accel(features)
        return ts

    # ---- copies / small accessors (svox.py:288-350, 600-645, 784-803) ------------------------------------------
    def snap(self, indices):
        """Lowest corner (world space) of the leaf that holds each point (svox.py:288-297)."""
        view = self[indices]
        per_leaf = view.corners                                   # one row per unique leaf, increasing slot order
        packed = self._pack_index(view.unique_leaf_node)
        return per_leaf[torch.searchsorted(packed, view.leaf_node_id)]

    def clone(self, device=None):
        """Deep copy, optionally onto another device (svox.py:299-350; the data channels of this fork are row indices,
        so the reference's channel selector of partial() has nothing to select and is not offered)."""
        dev = self.data.device if device is None else device
        t2 = N3Tree(N=self.N, data_dim=self.data_dim, depth_limit=self.depth_limit, geom_resize_fact=self.geom_resize_fact,
                    data_format=repr(self.data_format) if self.data_format is not None else None,
                    extra_data=self.extra_data.clone() if self.extra_data is not None else None, map_location=dev)
        for name in ("child", "data", "parent_depth", "invradius", "offset", "_n_internal", "_n_free"):
            setattr(t2, name, getattr(self, name).detach().clone().to(dev))
        t2.features = nn.Parameter(self.features.detach().clone().to(dev))
        t2.filled = self.filled
        t2._invalidate()
        return t2

    def partial(self, data_sel=None, device=None):
        if data_sel is not None:
            raise RuntimeError("svox_t_b200: partial(data_sel) selects float data channels, which this fork's index-valued "
                               "`data` tensor does not have; use clone()")
        return self.clone(device)

    def shrink_to_fit(self):
        """Drop the unused capacity of child / data / parent_depth (svox.py:600-643). Returns True if anything shrank.
        (This implementation never frees nodes, so there is no fragmentation to compact.)"""
        if self._lock_tree_structure:
            raise RuntimeError("Tree locked")
        n = self.filled
        if n >= self.capacity:
            return False
        self.child, self.data, self.parent_depth = (t[:n].contiguous() for t in (self.child, self.data, self.parent_depth))
        self._invalidate()
        return True

    @property
    def ndim(self):
        return 2

    @property
    def shape(self):
        return torch.Size((self.n_leaves, self.data_dim))

    def size(self, dim):
        return self.data_dim if dim == 1 else self.n_leaves

    def numel(self):
        return self.data_dim * self.n_leaves

    def accumulate_weights(self):
        """``with tree.accumulate_weights() as accum: render(...)`` then ``accum()`` -> per-leaf sum of the compositing
        weights of every ray rendered inside the block, in ``tree[:]`` order (svox.py:664-677, 948-970)."""
        return WeightAccumulator(self)

    def aux(self, arr):
        return self[:].aux(arr)

    # ---- persistence (svox.py:679-752) -----------------------------------------------------------------------
    def save(self, path, shrink=True, compress=True):
        n = self.filled if shrink else self.capacity
        blob = {
            "data_dim": self.data_dim, "child": self.child[:n].cpu().numpy(),
            "parent_depth": self.parent_depth[:n].cpu().numpy(), "n_internal": self.filled,
            "n_free": int(self._n_free.item()), "invradius3": self.invradius.cpu().numpy(),
            "offset": self.offset.cpu().numpy(), "depth_limit": self.depth_limit,
            "geom_resize_fact": self.geom_resize_fact, "data": self.data[:n].cpu().numpy(),
        }
        if self.data_format is not None:
            blob["data_format"] = repr(self.data_format)
        if self.extra_data is not None:
            blob["extra_data"] = self.extra_data.cpu().numpy()
        (np.savez_compressed if compress else np.savez)(path, **blob)

    @classmethod
    def load(cls, path, map_location="cpu"):
        z = np.load(path)
        extra = torch.from_numpy(z["extra_data"]).to(map_location) if "extra_data" in z.files else None
        tree = cls(extra_data=extra, map_location=map_location)
        tree.data_dim = int(z["data_dim"])
        tree.child = torch.from_numpy(z["child"]).to(map_location)
        tree.N = tree.child.shape[-1]
        tree.parent_depth = torch.from_numpy(z["parent_depth"]).to(map_location)
        tree._n_internal.fill_(int(z["n_internal"]))
        tree.filled = int(z["n_internal"])
        if "invradius3" in z.files:
            tree.invradius = torch.from_numpy(z["invradius3"].astype(np.float32)).to(map_location)
        else:
            tree.invradius.fill_(float(z["invradius"]))
        tree.offset = torch.from_numpy(z["offset"].astype(np.float32)).to(map_location)
        tree.depth_limit = int(z["depth_limit"])
        tree.geom_resize_fact = float(z["geom_resize_fact"])
        tree.data = torch.from_numpy(z["data"]).to(torch.int32).to(map_location)
        tree._n_free.fill_(int(z["n_free"]) if "n_free" in z.files else 0)
        tree.data_format = DataFormat(z["data_format"].item()) if "data_format" in z.files else None
        tree._invalidate()
        return tree

    # ---- magic ---------------------------------------------------------------------------------------------------
    def __repr__(self):
        return (f"svox_t_b200.N3Tree(N={self.N}, data_dim={self.data_dim}, depth_limit={self.depth_limit}, "
                f"capacity:{self.filled}/{self.capacity}, data_format:{self.data_format or 'RGBA'})")

    def __getitem__(self, key):
        return N3TreeView(self, key)

    def __setitem__(self, key, val):
        """``tree[pts] = values``: row assignment at world points (svox.py:767-768)."""
        if torch.is_tensor(key) and key.ndim == 2 and key.shape[1] == 3:
            val = torch.as_tensor(val, dtype=torch.float32, device=self.data.device)
            if val.ndim < 2:
                val = val.expand(key.shape[0], val.shape[0] if val.ndim == 1 else self.data_dim)
            self.set(key, val)
        else:
            N3TreeView(self, key).set(val)

    def __len__(self):
        return self.n_leaves


class WeightAccumulator:
    """svox.py:948-970. The tree structure is locked while weights accumulate."""

    def __init__(self, tree):
        self.tree = tree

    def __enter__(self):
        self.tree._lock_tree_structure = True
        self.tree._weight_accum = torch.zeros(self.tree.child.shape, dtype=torch.float32, device=self.tree.data.device)
        self.weight_accum = self.tree._weight_accum
        return self

    def __exit__(self, type, value, traceback):
        self.tree._weight_accum = None
        self.tree._lock_tree_structure = False

    @property
    def value(self):
        return self.weight_accum

    def __call__(self):
        return self.tree.aux(self.weight_accum)


def get_transformation_matrix(src_pose, tgt_pose):
    """svox.py:971-972."""
    return torch.matmul(tgt_pose, torch.inverse(src_pose))


def warp_vertices(transformation_matrix, coordinates, skinning_weights, joint_index):
    """Linear blend skinning of points: (coords'[P,3], mats[P,4,4]); svox.py:974-975."""
    return _WarpVerticalFunction.apply(transformation_matrix, coordinates, skinning_weights, joint_index)


def blend_transformation_matrix(transformation_matrix, skinning_weights, joint_index):
    """svox.py:978-981."""
    coordinates = torch.zeros((skinning_weights.size(0), 3), device=skinning_weights.device)
    _, matrices = _C.warp_vertices(transformation_matrix, coordinates, skinning_weights, joint_index)
    return matrices
