"""synth.py -- seeded synthetic workloads of the BASELINE.md shapes and the tree container.

Host-side utility shared by bench.py and the tests: random trees in Easel's convention, and alignments with a
nested random structure over ~60% of the columns, phylogenetically correlated rows, Beta(0.5,4) per-column gap
fractions (< 0.75, so --gapthresh drops nothing), 0.1% N and Gamma(2,1/2) weights normalised to sum N.
"""
import ctypes as C

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _TreeStruct(C.Structure):
    _fields_ = [("N", C.c_int), ("left", _ip), ("right", _ip), ("parent", _ip), ("ld", _dp), ("rd", _dp)]


class Tree:
    """Easel-convention binary tree held as numpy arrays."""

    def __init__(self, left, right, parent, ld, rd):
        self.left = np.ascontiguousarray(left, dtype=np.int32)
        self.right = np.ascontiguousarray(right, dtype=np.int32)
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.ld = np.ascontiguousarray(ld, dtype=np.float64)
        self.rd = np.ascontiguousarray(rd, dtype=np.float64)
        self.N = len(self.left) + 1

    def cstruct(self):
        return _TreeStruct(self.N, self.left.ctypes.data_as(_ip), self.right.ctypes.data_as(_ip), self.parent.ctypes.data_as(_ip),
                           self.ld.ctypes.data_as(_dp), self.rd.ctypes.data_as(_dp))


def random_tree(N, rng, mean_len=0.05):
    """Random binary tree with N leaves, parents numbered before children (preorder), exponential
    branch lengths.  Built by random splitting of leaf sets."""
    left = np.zeros(N - 1, np.int32)
    right = np.zeros(N - 1, np.int32)
    parent = np.zeros(N - 1, np.int32)
    ld = rng.exponential(mean_len, N - 1)
    rd = rng.exponential(mean_len, N - 1)
    leaves = rng.permutation(N)
    nxt = [1]
    # iterative preorder construction: stack of (node, lo, hi) over the permuted leaf array
    stack = [(0, 0, N)]
    while stack:
        v, lo, hi = stack.pop()
        n = hi - lo
        k = 1 if n == 2 else int(rng.integers(1, n))
        for side, (a, b) in enumerate(((lo, lo + k), (lo + k, hi))):
            if b - a == 1:
                child = -int(leaves[a])
            else:
                child = nxt[0]
                nxt[0] += 1
                parent[child] = v
                stack.append((child, a, b))
            if side == 0:
                left[v] = child
            else:
                right[v] = child
    # children were numbered in creation order, which is parent-before-child
    return Tree(left, right, parent, ld, rd)


def synthetic_msa(N, L, seed=42, gap_mean=0.11, frac_paired=0.6, n_frac=0.001, mean_len=0.05, weights="gamma"):
    """Seeded synthetic alignment of the SURVEY 8d recipe: random nested structure over ~60% of the
    columns, tree-evolved columns (HKY-like 4x4 for unpaired, pair-preserving moves for paired),
    Beta(0.5,4)-distributed per-column gap fractions (< 0.75), 0.1% N, weights Gamma(2,1/2)
    normalised to sum N.  Returns (ax uint8 [N][L], wgt float64 [N], pair partner array)."""
    rng = np.random.default_rng(seed)
    # nested structure: random helices
    partner = -np.ones(L, int)
    target = int(frac_paired * L) // 2
    tries = 0
    while (partner >= 0).sum() // 2 < target and tries < 20 * L:
        tries += 1
        hl = int(rng.integers(3, 9))
        i = int(rng.integers(0, max(1, L - 2 * hl - 4)))
        lo, hi = i + 2 * hl + 3, min(L, i + 2 * hl + 4 + max(4, L // 4))
        if lo >= hi:
            continue
        j = int(rng.integers(lo, hi))
        if j >= L:
            continue
        a = np.arange(i, i + hl)
        b = j - np.arange(hl)
        span = np.arange(i, j + 1)
        if (partner[a] >= 0).any() or (partner[b] >= 0).any():
            continue
        inner = partner[span]
        inner = inner[inner >= 0]
        if ((inner < i) | (inner > j)).any():
            continue                                               # would cross an existing helix
        partner[a] = b
        partner[b] = a
    # evolve down a random "caterpillar-ish" tree implicitly: sequence s copies a random earlier
    # sequence and mutates; cheap O(N L), gives phylogenetic correlation
    ax = np.empty((N, L), dtype=np.uint8)
    ax[0] = rng.integers(0, 4, L)
    wc = {0: 3, 3: 0, 1: 2, 2: 1}
    for c in range(L):
        if partner[c] > c:
            ax[0, partner[c]] = wc[int(ax[0, c])]
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    for s in range(1, N):
        par = int(rng.integers(max(0, s - 50), s))
        row = ax[par].copy()
        mut = rng.random(L) < mean_len * rng.exponential(1.0) * 3
        new = rng.integers(0, 4, L).astype(np.uint8)
        row = np.where(mut, new, row)
        # keep pairs complementary 90% of the time
        up = np.nonzero((partner > np.arange(L)) & (mut | mut[np.maximum(partner, 0)]))[0]
        keep = rng.random(len(up)) < 0.9
        row[partner[up[keep]]] = comp[row[up[keep]]]
        ax[s] = row
    gapf = np.minimum(rng.beta(0.5, 4.0, L) * (gap_mean / 0.111), 0.7)
    gaps = rng.random((N, L)) < gapf[None, :]
    ax[gaps] = 4
    ax[rng.random((N, L)) < n_frac] = 15
    if weights == "gamma":
        w = rng.gamma(2.0, 0.5, N)
        w *= N / w.sum()
    elif weights == "ones":
        w = np.ones(N)
    else:
        raise ValueError("weights must be 'gamma' or 'ones' (PB / GSC weights are test-side preprocessing)")
    return ax, w, partner


def synthetic_family(N, L, seed=42, mean_len=0.01, gap_mean=0.11, frac_paired=0.6, n_frac=0.001, weights="gamma"):
    """Seeded synthetic alignment TOGETHER WITH the tree it evolved on (R-scape infers its tree from the alignment, so
    the null generators always see a tree that explains the data; a tree unrelated to the alignment would make the
    parsimony reconstruction place substitutions on nearly every branch and column).  Same recipe as synthetic_msa
    otherwise: nested random helices over ~60% of the columns kept complementary 90% of the time, Beta(0.5,4)
    per-column gap fractions, 0.1% N, Gamma(2,1/2) weights.  Branch lengths are Exp(mean_len) substitutions per site.
    Returns (ax uint8 [N][L], wgt float64 [N], pair partner array, Tree)."""
    rng = np.random.default_rng(seed)
    tree = random_tree(N, rng, mean_len)
    partner = -np.ones(L, int)
    target = int(frac_paired * L) // 2
    tries = 0
    while (partner >= 0).sum() // 2 < target and tries < 20 * L:
        tries += 1
        hl = int(rng.integers(3, 9))
        i = int(rng.integers(0, max(1, L - 2 * hl - 4)))
        lo, hi = i + 2 * hl + 3, min(L, i + 2 * hl + 4 + max(4, L // 4))
        if lo >= hi:
            continue
        j = int(rng.integers(lo, hi))
        if j >= L:
            continue
        a = np.arange(i, i + hl)
        b = j - np.arange(hl)
        span = np.arange(i, j + 1)
        if (partner[a] >= 0).any() or (partner[b] >= 0).any():
            continue
        inner = partner[span]
        inner = inner[inner >= 0]
        if ((inner < i) | (inner > j)).any():
            continue
        partner[a] = b
        partner[b] = a
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    upper = np.nonzero(partner > np.arange(L))[0]
    root = rng.integers(0, 4, L).astype(np.uint8)
    root[partner[upper]] = comp[root[upper]]
    anc = np.empty((max(N - 1, 1), L), dtype=np.uint8)
    anc[0] = root
    ax = np.empty((N, L), dtype=np.uint8)
    for v in range(N - 1):                                      # parents are numbered before their children
        for child, t in ((tree.left[v], tree.ld[v]), (tree.right[v], tree.rd[v])):
            row = anc[v].copy()
            mut = np.nonzero(rng.random(L) < 1.0 - np.exp(-t))[0]
            if len(mut):
                row[mut] = rng.integers(0, 4, len(mut))
                pm = mut[partner[mut] >= 0]
                keep = pm[rng.random(len(pm)) < 0.9]
                lo_side = np.minimum(keep, partner[keep])
                row[partner[lo_side]] = comp[row[lo_side]]
            if child > 0:
                anc[child] = row
            else:
                ax[-child] = row
    gapf = np.minimum(rng.beta(0.5, 4.0, L) * (gap_mean / 0.111), 0.7)
    ax[rng.random((N, L)) < gapf[None, :]] = 4
    ax[rng.random((N, L)) < n_frac] = 15
    if weights == "gamma":
        w = rng.gamma(2.0, 0.5, N)
        w *= N / w.sum()
    elif weights == "ones":
        w = np.ones(N)
    else:
        raise ValueError("weights must be 'gamma' or 'ones'")
    return ax, w, partner, tree

