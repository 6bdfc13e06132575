"""Deterministic synthetic scenes for the octree volume-rendering hot path.

Everything here is host-side numpy/torch-CPU: octrees emitted directly in the reference tensor format
(``child[n,2,2,2]`` relative offsets, ``data[n,2,2,2,1]`` feature-row indices or the empty sentinel,
``parent_depth[n,2]``; svox_t/svox.py:121-139, 538-543), leaf feature tables, ray batches and look-at
cameras. Used by tests/, bench.py and __graft_entry__.smoke(); identical inputs can be fed to the
reference extension because it consumes the same tensors (svox_t/svox.py:726-739 assigns them on load).
"""
import math

import numpy as np

SENTINEL = 1410065408  # int(1e10) wrapped to int32 (svox_t/svox.py:124)


def _occupied_keys(L, shape, r_out=0.30, r_in=0.27, chunk=32):
    """Sorted int64 keys (i*R+j)*R+k of the occupied finest voxels (centres (i+.5)/R) at resolution R=2^L."""
    R = 1 << L
    if shape == "all":
        return np.arange(R ** 3, dtype=np.int64)
    c = (np.arange(R, dtype=np.float64) + 0.5) / R - 0.5
    c2 = c * c
    yz2 = c2[:, None] + c2[None, :]
    keys = []
    for i0 in range(0, R, chunk):
        i1 = min(R, i0 + chunk)
        d2 = c2[i0:i1, None, None] + yz2[None]
        if shape == "ball":
            m = d2 < r_out * r_out
        elif shape == "shell":
            m = (d2 < r_out * r_out) & (d2 > r_in * r_in)
        else:
            raise ValueError(shape)
        ii, jj, kk = np.nonzero(m)
        keys.append(((ii.astype(np.int64) + i0) * R + jj) * R + kk)
    return np.concatenate(keys) if keys else np.zeros(0, np.int64)


def _unkey(keys, R):
    k = keys % R
    j = (keys // R) % R
    i = keys // (R * R)
    return i, j, k


def tree_from_voxels(keys, L):
    """Octree with one finest-level (depth L) leaf per occupied voxel key and coarse empty leaves elsewhere.

    Internal node for every occupied cell at levels 0..L-1, numbered BFS by level and, within a level, by
    the key (i*r+j)*r+k. ``data`` = rank of the voxel in sorted key order, or SENTINEL.
    Returns dict(child, data, parent_depth, M, n_nodes, n_leaves, L).
    """
    assert L >= 1
    keys = np.unique(np.asarray(keys, dtype=np.int64))
    M = int(keys.shape[0])
    # cells[l] = sorted unique keys of occupied cells at level l (resolution 2^l), l = 0..L
    cells = [None] * (L + 1)
    cells[L] = keys
    for l in range(L, 0, -1):
        R = 1 << l
        i, j, k = _unkey(cells[l], R)
        Rp = R >> 1
        cells[l - 1] = np.unique(((i >> 1) * Rp + (j >> 1)) * Rp + (k >> 1))
    if M == 0:
        cells[0] = np.zeros(1, np.int64)
    base = np.zeros(L + 1, dtype=np.int64)
    for l in range(1, L + 1):
        base[l] = base[l - 1] + (len(cells[l - 1]) if l - 1 < L else 0)
    n_nodes = int(base[L - 1] + len(cells[L - 1])) if L >= 1 else 1
    child = np.zeros((n_nodes, 2, 2, 2), dtype=np.int32)
    data = np.full((n_nodes, 2, 2, 2, 1), SENTINEL, dtype=np.int32)
    parent_depth = np.zeros((n_nodes, 2), dtype=np.int32)
    for l in range(1, L + 1):
        R = 1 << l
        Rp = R >> 1
        ck = cells[l]
        if len(ck) == 0:
            continue
        i, j, k = _unkey(ck, R)
        pkey = ((i >> 1) * Rp + (j >> 1)) * Rp + (k >> 1)
        prank = np.searchsorted(cells[l - 1], pkey)
        pnode = base[l - 1] + prank
        u, v, w = (i & 1), (j & 1), (k & 1)
        if l < L:
            node = base[l] + np.arange(len(ck), dtype=np.int64)
            child[pnode, u, v, w] = (node - pnode).astype(np.int32)
            parent_depth[node, 0] = (pnode * 8 + u * 4 + v * 2 + w).astype(np.int32)
            parent_depth[node, 1] = l
        else:
            data[pnode, u, v, w, 0] = np.arange(len(ck), dtype=np.int32)
    n_leaves = int((child == 0).sum())
    return dict(child=child, data=data, parent_depth=parent_depth, M=M, n_nodes=n_nodes,
                n_leaves=n_leaves, L=L, N=2)


def synth_tree(L, shape="ball", r_out=0.30, r_in=0.27):
    """C1: synth_tree(4, 'all'); C2/C3/C4: synth_tree(8, 'ball'); C5: synth_tree(10, 'shell')."""
    return tree_from_voxels(_occupied_keys(L, shape, r_out, r_in), L)


def voxel_centers(keys, L):
    R = 1 << L
    i, j, k = _unkey(np.asarray(keys, dtype=np.int64), R)
    return ((np.stack([i, j, k], -1).astype(np.float64) + 0.5) / R).astype(np.float32)


def synth_features(M, D, seed=0):
    """Channels 0..D-2 ~ N(0,1); sigma channel ~ U(-2, 8) (about 20 % of rows fail the sigma > 0 test)."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    f = torch.randn(M, D, generator=g, dtype=torch.float32)
    f[:, D - 1] = torch.rand(M, generator=g, dtype=torch.float32) * 10.0 - 2.0
    return f.numpy()


def _unit(rng, n):
    v = rng.standard_normal((n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def synth_rays(Q, seed=1, center=0.5, r_origin=1.5, r_target=0.30):
    """Origins uniform on the sphere |o-c| = 1.5, targets uniform in the ball radius 0.30; every ray hits the cube."""
    rng = np.random.default_rng(seed)
    o = center + r_origin * _unit(rng, Q)
    tgt = center + r_target * _unit(rng, Q) * np.cbrt(rng.random((Q, 1)))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


def fibonacci_dirs(n):
    """n roughly uniform unit vectors; the single-view case is the SURVEY's (0.4, 0.5, 0.766) direction."""
    if n == 1:
        u = np.array([[0.4, 0.5, 0.766]])
        return u / np.linalg.norm(u)
    k = np.arange(n) + 0.5
    y = 1.0 - 2.0 * k / n
    r = np.sqrt(np.maximum(0.0, 1.0 - y * y))
    phi = k * math.pi * (3.0 - math.sqrt(5.0))
    return np.stack([r * np.cos(phi), y, r * np.sin(phi)], -1)


def look_at(eye, target=(0.5, 0.5, 0.5), up=(0.0, 1.0, 0.0)):
    """OpenGL-style c2w [4,4] float32: camera looks down -z, +y up (svox_t/csrc/rt_kernel.cu:1158-1165)."""
    eye, target, up = (np.asarray(a, dtype=np.float64) for a in (eye, target, up))
    z = eye - target
    z /= np.linalg.norm(z)
    x = np.cross(up, z)
    if np.linalg.norm(x) < 1e-8:
        x = np.cross(np.array([1.0, 0.0, 0.0]), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, eye
    return c2w.astype(np.float32)


def synth_cameras(n_views=1, dist=1.0, center=0.5):
    return [look_at(center + dist * u) for u in fibonacci_dirs(n_views)]


def synth_skeleton(P, J=24, B=4, seed=3):
    """Skinning weights (Dirichlet) + joint indices for P points, and per-joint rigid transforms [J,4,4]."""
    rng = np.random.default_rng(seed)
    w = rng.dirichlet(np.ones(B), size=P).astype(np.float32)
    ji = rng.integers(0, J, size=(P, B)).astype(np.int32)
    rng4 = np.random.default_rng(seed + 1)
    axis = _unit(rng4, J)
    ang = np.deg2rad(15.0) * rng4.random(J)
    T = np.zeros((J, 4, 4), dtype=np.float64)
    for j in range(J):
        a, (x, y, z) = ang[j], axis[j]
        K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
        Rm = np.eye(3) + math.sin(a) * K + (1 - math.cos(a)) * (K @ K)
        tr = 0.02 * (2 * rng4.random(3) - 1)
        # rotate about the scene centre so warped points stay inside the unit cube
        c = np.full(3, 0.5)
        T[j, :3, :3] = Rm
        T[j, :3, 3] = c - Rm @ c + tr
        T[j, 3, 3] = 1.0
    return T.astype(np.float32), w, ji
