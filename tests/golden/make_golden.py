#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference tree (run in the build container only).

  tests/golden/arisong_tutorial.npz   the tutorial alignment tutorial/updated_Arisong.sto as digital residues
                                      (uint8 95x150, Easel RNA codes) + SS_cons
  tests/golden/arisong_tutorial.json  the known answers printed in documentation/tutorial.tex:187-212 for
                                      `R-scape -s tutorial/updated_Arisong.sto`: banner numbers and the 11
                                      significant pairs with their GTp scores (5 decimals)
  tests/golden/ref_scans.npz          outputs of the REFERENCE's own correlators.c (oracle/_ref) on seeded synthetic
                                      alignments: every statistic x class x correction on one small alignment, used
                                      to pin the oracle wherever oracle/_ref is not available (e.g. the GPU box)

  tests/golden/arisong_fasttree.nwk   FastTree 2.1.11 (the reference's vendored copy, `-quiet -nt`) on the analysed tutorial alignment
  tests/golden/arisong_fasttree.npz   ... read, re-ordered to the alignment's rows and rooted at the midpoint by the reference's own code
  tests/golden/ref_evalues.npz        the reference's static cov2evalue / evalue2cov on a seeded null histogram (with / without a tail)
  tests/golden/ref_treesubs.npz       the reference's Tree_Substitutions on a seeded alignment + tree

Usage: python tests/golden/make_golden.py   (needs /root/reference and a built oracle/_ref)
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

REF = "/root/reference"


def make_fasttree_fixture(po, names, ax):
    """tests/golden/arisong_fasttree.{nwk,npz}: the tree R-scape's default null model is built on for the tutorial alignment
    (Tree_CalculateExtFromMSA, src/msatree.c:49-105): the analysed alignment (gap columns removed) written as aligned FASTA, the
    reference's vendored FastTree 2.1.11 run as `FastTree -quiet -nt` (:149), its Newick output read by the shim's
    esl_tree_ReadNewick and re-ordered / rooted at the midpoint by the reference's own Tree_ReorderTaxaAccordingMSA +
    Tree_RootAtMidPoint (oracle/_ref).  FastTree is deterministic, so the fixture is reproducible from the reference tree."""
    import subprocess
    import tempfile
    sub, keep = po.remove_gap_columns(ax)
    sub = po.degen_to_N(sub)
    text = {0: "A", 1: "C", 2: "G", 3: "U", 4: "-", 15: "N"}
    fasttree = os.path.join(ROOT, "oracle", "_ref", "FastTree")
    with tempfile.TemporaryDirectory() as tmp:
        afa = os.path.join(tmp, "msa.afa")
        with open(afa, "w") as fh:
            for name, row in zip(names, sub):
                fh.write(">" + name + "\n" + "".join(text[int(c)] for c in row) + "\n")
        nwk = subprocess.run([fasttree, "-quiet", "-nt", afa], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    path = os.path.join(HERE, "arisong_fasttree.nwk")
    with open(path, "w") as fh:
        fh.write(nwk)
    tree = po.RefLib().tree_from_newick(path, names, rootatmid=True)
    np.savez_compressed(os.path.join(HERE, "arisong_fasttree.npz"), left=tree.left, right=tree.right, parent=tree.parent, ld=tree.ld, rd=tree.rd,
                        names=np.array(names))
    return tree


def main():
    po = ge.load_oracle()
    names, ax, ss = po.read_stockholm(os.path.join(REF, "tutorial", "updated_Arisong.sto"))
    np.savez_compressed(os.path.join(HERE, "arisong_tutorial.npz"), ax=ax, ss_cons=np.array(ss))

    tex = open(os.path.join(REF, "documentation", "tutorial.tex")).read().splitlines()
    block = tex[186:214]                                   # the `-s` transcript, tutorial.tex:187-214
    banner = next(l for l in block if l.startswith("# MSA updated_Arisong_1"))
    m = re.search(r"nseq (\d+) \((\d+)\) alen (\d+) \((\d+)\) avgid ([\d.]+) \(([\d.]+)\) nbpairs (\d+)", banner)
    summary = next(l for l in block if l.startswith("# GTp"))
    pairs = []
    for l in block:
        if l.startswith("*"):
            f = l.split()
            pairs.append(dict(i=int(f[1]), j=int(f[2]), score=float(f[3]), evalue=float(f[4]), pvalue=float(f[5])))
    gold = dict(source="documentation/tutorial.tex:187-212", nseq=int(m.group(1)), alen=int(m.group(3)), alen_orig=int(m.group(4)),
                avgid=float(m.group(5)), nbpairs=int(m.group(7)), summary=summary, pairs=pairs)
    with open(os.path.join(HERE, "arisong_tutorial.json"), "w") as fh:
        json.dump(gold, fh, indent=1)

    make_fasttree_fixture(po, names, ax)

    ref = po.RefLib()
    msa, wgt, _ = po.synthetic_msa(90, 58, seed=2024)
    out = dict(msa=msa, wgt=wgt)
    for stat in ("GT", "CHI", "MI", "MIr", "MIg", "OMES", "CCF", "RAF", "RAFS"):
        for cls in ("C16", "C2", "CWC"):
            if cls == "CWC" and stat != "GT":
                continue
            for ac in ("APC", "ASC", "NOCORR"):
                r = ref.scan(msa, wgt, getattr(po, stat), getattr(po, cls), getattr(po, ac))
                out[f"{stat}_{cls}_{ac}_cov"] = r["cov"]
                out[f"{stat}_{cls}_{ac}_mm"] = np.array([r["mincov"], r["maxcov"]])
    r = ref.scan(msa, wgt, po.GT, po.C16, po.APC)
    for k in ("pp", "pm", "ps", "nseff", "ngap"):
        out["probs_" + k] = r[k]
    np.savez_compressed(os.path.join(HERE, "ref_scans.npz"), **out)

    # E-values: the reference's own static cov2evalue / evalue2cov (src/covariation.c:2370-2435, reached through
    # oracle/ref_glue_evalue.c) on a seeded null histogram, without and with a fitted tail
    rng = np.random.default_rng(7)
    x = np.maximum(rng.gamma(2.0, 2.5, 100000) - 8.0, -10 + 0.05)
    b = np.ceil((x + 10) / 0.05 - 1).astype(np.int64)
    obs = np.bincount(b, minlength=int(b.max()) + 6).astype(np.uint64)
    plain = po.NullFit(-10.0, 0.05, obs, xmax=float(x.max()))
    fit = plain.exp_tail(0.05)
    scores = np.concatenate([rng.uniform(-13, plain.bmin + plain.w * (2 * plain.nb + 4), 600), plain.bmin + plain.w * np.arange(0, 2 * plain.nb + 2, 7)])
    ev = dict(obs=obs, xmax=plain.xmax, phi=fit.phi, cmin=fit.cmin, survfit=fit.survfit, scores=scores,
              thresholds=np.array([1e-6, 1e-3, 0.05, 1.0, 10.0, 1e4]))
    for name, null in (("plain", plain), ("fit", fit)):
        for Nc in (1, 1225):
            ev[f"{name}_cov2evalue_{Nc}"] = np.array([ref.cov2evalue(v, null, Nc) for v in scores])
            ev[f"{name}_evalue2cov_{Nc}"] = np.array([ref.evalue2cov(e, null, Nc) for e in ev["thresholds"]])
    np.savez_compressed(os.path.join(HERE, "ref_evalues.npz"), **ev)

    # Tree_Substitutions (src/msatree.c:1423-1554, its own Fitch pass on the shim's MT19937 stream)
    msa = po.synthetic_msa(36, 40, seed=77)[0]
    tree = po.random_tree(36, np.random.default_rng(77))
    ts = dict(msa=msa, left=tree.left, right=tree.right, parent=tree.parent, ld=tree.ld, rd=tree.rd, seed=77)
    for g in (0, 1):
        ns, nd, nj = ref.tree_substitutions(77, tree, msa, bool(g))
        ts[f"nsubs_{g}"], ts[f"ndouble_{g}"], ts[f"njoin_{g}"] = ns, nd, nj
    np.savez_compressed(os.path.join(HERE, "ref_treesubs.npz"), **ts)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
