"""BASELINE config 1 pieces shared by the CPU and GPU end-to-end tests: `R-scape -s tutorial/updated_Arisong.sto` with R-scape's
defaults (GTp, APC, 20 tree-shuffled nulls on the FastTree tree, gamma tail fit, E < 0.05, two-set test), against the transcript
documentation/tutorial.tex:187-212 (tests/golden/arisong_tutorial.json)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BMIN, W0, HPTS, TOL = -10.0, 0.05, 400, 1e-6          # BMIN, cfg->w, HPTS, tol (src/covariation.h:22, src/R-scape.c:426)
PMASS, FRACFIT, ETHRESH, NSHUFFLE = 0.0005, 1.0, 0.05, 20   # --pmass, --fracfit, -E, nshuffle for nseq > 40 (src/R-scape.c:428-429, 2451)


def wuss_pairs(ss):
    """Base pairs (0-based columns) of a WUSS string: nested brackets <> () [] {}."""
    close = {">": "<", ")": "(", "]": "[", "}": "{"}
    stacks, pairs = {o: [] for o in close.values()}, []
    for k, c in enumerate(ss):
        if c in stacks:
            stacks[c].append(k)
        elif c in close:
            pairs.append((stacks[close[c]].pop(), k))
    return sorted(pairs)


def load(po):
    """-> analysed alignment (gap columns removed, degenerate -> N), GSC weights, kept columns, structure mask, tree, transcript."""
    z = np.load(os.path.join(GOLD, "arisong_tutorial.npz"))
    with open(os.path.join(GOLD, "arisong_tutorial.json")) as fh:
        gold = json.load(fh)
    sub, keep = po.remove_gap_columns(z["ax"])              # msamanip_RemoveGapColumns, --gapthresh 0.75
    sub = po.degen_to_N(sub)
    wgt = po.weights_gsc(sub)                               # nseq <= 1000: esl_msaweight_GSC (src/R-scape.c:1555)
    col = {int(c): k for k, c in enumerate(keep)}
    L = sub.shape[1]
    mask = np.zeros((L, L), np.uint8)
    for i, j in wuss_pairs(str(z["ss_cons"])):
        if i in col and j in col:
            mask[col[i], col[j]] = 1
    t = np.load(os.path.join(GOLD, "arisong_fasttree.npz"))
    tree = po.Tree(t["left"], t["right"], t["parent"], t["ld"], t["rd"])
    return sub, wgt, keep, mask, tree, gold


def null_width(lo, hi):
    """calculate_width_histo, src/R-scape.c:1355-1360"""
    w = min(W0, (hi - max(BMIN, lo)) / HPTS)
    return 0.0 if w < TOL else w


def null_bins_needed(w, xmax_null, maxcov_input):
    """bins covering the nulls and the input alignment's own histogram (bmax = maxCOV + 5 w, src/covariation.c:415, 461-466)"""
    return int(np.ceil((max(xmax_null, maxcov_input) + 5 * w - BMIN) / w)) + 1


def called_pairs(hits, keep):
    """hit list -> set of (i, j) in the 1-based coordinates of the input alignment, as the transcript prints them"""
    return {(int(keep[i]) + 1, int(keep[j]) + 1) for i, j in zip(hits["i"], hits["j"])}
