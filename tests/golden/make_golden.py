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
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
