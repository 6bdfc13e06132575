"""Small host-side helpers of the N3Tree API: ``N3TreeView`` (the ``tree[points]`` selector used by the per-frame
rebuild), ``LocalIndex`` and ``DataFormat``. Behaviour follows the reference (svox_t/helpers.py:38-109, 378-420);
the legacy value accessors that index ``tree.data`` as floats (helpers.py:111-338) are out of scope.
"""
import torch


class LocalIndex:
    """Query with points already in tree space [0,1]^3: ``tree[LocalIndex(pts)]`` (helpers.py:378-384)."""

    def __init__(self, val):
        self.val = val


class DataFormat:
    """Parsed ``data_format`` string: "RGBA", "SH9", "SG25", "ASG4" ... (helpers.py:386-420)."""
    RGBA, SH, SG, ASG = 0, 1, 2, 3
    _NAMES = {0: "RGBA", 1: "SH", 2: "SG", 3: "ASG"}

    def __init__(self, txt):
        head = txt.rstrip("0123456789")
        tail = txt[len(head):]
        self.format = {"SH": self.SH, "SG": self.SG, "ASG": self.ASG}.get(head, self.RGBA) if tail else self.RGBA
        self.basis_dim = int(tail) if tail else -1

    def __repr__(self):
        return self._NAMES[self.format] + (str(self.basis_dim) if self.basis_dim >= 0 else "")


class N3TreeView:
    """Selection of leaves: by world points (``tree[pts]``), tree-space points (``tree[LocalIndex(pts)]``) or a
    slice over all leaves (``tree[:]``). ``.refine()`` splits the selected unique leaves (helpers.py:101-109)."""

    def __init__(self, tree, key):
        self.tree = tree
        local = False
        if isinstance(key, LocalIndex):
            key, local = key.val, True
        if isinstance(key, tuple) and len(key) >= 3 and not torch.is_tensor(key[0]):
            key = torch.tensor(key[:3], dtype=torch.float32, device=tree.data.device).reshape(1, 3)
        if torch.is_tensor(key) and key.ndim == 2 and key.shape[1] == 3:
            pts = key if key.dtype == torch.float32 else key.float()
            _, node_ids, leaf_node = tree.forward(tree.features, pts.contiguous(), want_node_ids=True,
                                                  world=not local, want_leaf_node=True)
            self._packed_ids = node_ids          # packed slot id of every query point
            self.leaf_node_id = node_ids
            self.unique_leaf_node = leaf_node    # [n_hit, 4] unique leaves, increasing slot order
        else:
            self._packed_ids = None
            self.leaf_node_id = None
            leaves = tree._all_leaves()
            if isinstance(key, int):
                key = slice(key, key + 1)
            self.unique_leaf_node = leaves[key]
        self.key = (*self.unique_leaf_node.T,)
        self._tree_ver = tree._ver

    def _check_ver(self):
        if self.tree._ver > self._tree_ver:
            raise RuntimeError("N3TreeView has been invalidated because tree data layout has changed")

    def refine(self, repeats=1):
        self._check_ver()
        return self.tree.refine(repeats, sel=self.key, leaf_node=self.unique_leaf_node)

    @property
    def depths(self):
        """Depth of each selected leaf's node (root = 0)."""
        self._check_ver()
        return self.tree.parent_depth[self.key[0], 1]

    @property
    def lengths_local(self):
        """Side length of each selected leaf in tree space."""
        return float(self.tree.N) ** (-self.depths.float() - 1.0)

    @property
    def lengths(self):
        return self.lengths_local[:, None] / self.tree.invradius

    @property
    def corners_local(self):
        """Lower corner of each selected leaf in tree space (svox.py:804-826 -> calc_corners)."""
        self._check_ver()
        return self.tree._calc_corners(self.unique_leaf_node)

    @property
    def corners(self):
        return self.tree.tree2world(self.corners_local)

    def _rows(self):
        self._check_ver()
        idx = self.tree.data[self.key][..., 0].long()
        return idx, idx < self.tree.features.shape[0]

    @property
    def values(self):
        """Feature rows of the selected leaves, (n_leaves, data_dim), autograd enabled; zeros for empty leaves. (The
        reference's accessor, helpers.py:111-120, still indexes ``tree.data`` as if it held the floats.)"""
        idx, valid = self._rows()
        f = self.tree.features
        out = f.new_zeros((idx.shape[0], f.shape[1]))
        out[valid] = f[idx[valid]]
        return out

    @property
    def values_nograd(self):
        with torch.no_grad():
            return self.values

    def set(self, value):
        """Overwrite the feature rows of the selected leaves (helpers.py:267-270); empty leaves are skipped."""
        idx, valid = self._rows()
        f = self.tree.features
        value = torch.as_tensor(value, dtype=f.dtype, device=f.device)
        if value.ndim == 2 and value.shape[0] == idx.shape[0]:
            value = value[valid]
        with torch.no_grad():
            f[idx[valid]] = value

    # In-place element-wise updates of the selected leaves' feature rows (helpers.py:246-306; the reference still applies
    # them to ``tree.data``, which holds row indices in this fork). Empty leaves have no row and are skipped.
    def _apply_(self, fn):
        idx, valid = self._rows()
        f = self.tree.features
        rows = idx[valid]
        with torch.no_grad():
            f[rows] = fn(f[rows])

    def normal_(self, mean=0.0, std=1.0):
        self._apply_(lambda v: torch.randn_like(v) * std + mean)

    def uniform_(self, min=0.0, max=1.0):
        self._apply_(lambda v: torch.rand_like(v) * (max - min) + min)

    def clamp_(self, min=None, max=None):
        self._apply_(lambda v: v.clamp(min, max))

    def relu_(self):
        self._apply_(torch.relu)

    def sigmoid_(self):
        self._apply_(torch.sigmoid)

    def nan_to_num_(self, inf_val=2e4):
        self._apply_(lambda v: torch.nan_to_num(v, nan=0.0, posinf=inf_val, neginf=-inf_val))

    @property
    def shape(self):
        self._check_ver()
        return torch.Size((len(self), self.tree.features.shape[1]))

    @property
    def ndim(self):
        return 2

    def _indexer(self):
        return torch.stack(self.key[:4], dim=-1)

    def __repr__(self):
        return f"N3TreeView({len(self)} leaves of {self.tree!r})"

    def aux(self, arr):
        """Index an auxiliary per-slot array of shape (capacity, N, N, N, ...) with this view (helpers.py:239-244)."""
        self._check_ver()
        return arr[self.key]

    def __len__(self):
        return self.unique_leaf_node.shape[0]
