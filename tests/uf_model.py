"""The parallel (union-find) reformulation of the reference's seeded BFS, as a slow numpy/scipy model.

SURVEY.md A.4: the Labeler only ever writes label 1, so Alfa_GS::EraseBFS's label image (reference
src/alfa_gs.cpp:717-739 + Labeler/label.cpp:48-104) is pure reachability:
  1. components of the NON-EMPTY pixels under 4-neighbour edges (column wrap, no row wrap) with
     |angle difference| <= K;
  2. a component is ground iff it holds an accepted seed that is non-empty;
  3. L = pixels of ground components  U  accepted seeds that are empty  U  empty pixels 4-adjacent to a
     ground pixel u with |angle(u) - angle(empty)| <= K.
This is the formulation the CUDA kernels implement; tests check it against the BFS oracle on the CPU.
"""
import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def components(valid, angle, K):
    """-> int32 (H, W): smallest flat index of each non-empty pixel's component, -1 for empty pixels."""
    H, W = valid.shape
    idx = np.arange(H * W).reshape(H, W)
    a = angle.astype(np.int64)
    src, dst = [], []
    ok = valid[1:, :] & valid[:-1, :] & (np.abs(a[1:, :] - a[:-1, :]) <= K)
    src.append(idx[1:, :][ok]); dst.append(idx[:-1, :][ok])
    right = np.roll(idx, -1, axis=1)
    ok = valid & np.roll(valid, -1, axis=1) & (np.abs(a - np.roll(a, -1, axis=1)) <= K)
    if W > 1:
        src.append(idx[ok]); dst.append(right[ok])
    src = np.concatenate(src); dst = np.concatenate(dst)
    g = coo_matrix((np.ones(len(src), np.int8), (src, dst)), shape=(H * W, H * W))
    _, comp = connected_components(g, directed=False)
    first = np.full(comp.max() + 1, H * W, np.int64)
    flat_valid = valid.reshape(-1)
    np.minimum.at(first, comp[flat_valid], np.arange(H * W)[flat_valid])
    out = np.where(flat_valid, first[comp], -1).astype(np.int32)
    return out.reshape(H, W)


def seeds(valid):
    """Seed row per column (src/alfa_gs.cpp:722-725): bottom-most non-empty row, stopping at row 0."""
    H, W = valid.shape
    rows = np.zeros(W, np.int64)
    for c in range(W):
        r = H - 1
        while r > 0 and not valid[r, c]:
            r -= 1
        rows[c] = r
    return rows


def ground_labels(range_image, angle, K, S):
    valid = range_image != 0
    H, W = valid.shape
    comp = components(valid, angle, K)
    sr = seeds(valid)
    cols = np.arange(W)
    accepted = angle[sr, cols] <= S
    ground_roots = set(comp[sr[accepted], cols[accepted]].tolist()) - {-1}
    Lv = valid & np.isin(comp, list(ground_roots))
    L = Lv.copy()
    # accepted seeds that are empty pixels get the label but do not expand
    for c in cols[accepted]:
        if not valid[sr[c], c]:
            L[sr[c], c] = True
    a = angle.astype(np.int64)
    for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        # neighbour u = (r+dr, c+dc) of the empty pixel (r, c)
        u_L = np.roll(Lv, (-dr, -dc), axis=(0, 1))
        u_a = np.roll(a, (-dr, -dc), axis=(0, 1))
        ok = ~valid & u_L & (np.abs(u_a - a) <= K)
        if dr == -1:
            ok[0, :] = False  # no row wrap
        if dr == 1:
            ok[-1, :] = False
        L |= ok
    return L.astype(np.uint8)
