"""struct mutual_s / struct data_s of include/rscape_compat.h must be layout-identical to the reference's
src/correlators.h (the drop-in boundary).  Compiles two tiny programs, one against each header, and compares
sizeof / offsetof.  Needs the reference tree (skipped on the GPU box)."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FIELDS_MI = ["alen", "nseq", "pp", "pm", "nseff", "ps", "ngap", "type", "COV", "Eval", "besthreshCOV", "minCOV", "maxCOV", "ishuffled",
             "nseqthresh", "alenthresh", "abc"]
FIELDS_DATA = ["ofile", "r", "samplesize", "ranklist_null", "mi", "pt", "thresh", "statsmethod", "covmethod", "mode", "covtype", "OL", "nseq",
               "ctlist", "spair", "power", "r3d", "agg_Eval", "agg_method", "T", "ribosum", "ct", "clist", "msa2pdb", "msamap", "firstpos", "bmin",
               "w", "fracfit", "pmass", "doexpfit", "tau", "mu", "lambda", "allowpair", "tol", "nofigures", "verbose", "errbuf", "prep_RF"]

FIELDS_HIT = ["i", "j", "sc", "Eval", "pval", "nsubs", "power", "bptype", "is_compatible"]
FIELDS_HITLIST = ["nhit", "srthit", "hit", "Nt", "Nb"]

PROG = r'''
#include <stdio.h>
#include <stddef.h>
%s
int main(void) {
  printf("sizeof mutual_s %%zu\n", sizeof(struct mutual_s));
  printf("sizeof data_s %%zu\n", sizeof(struct data_s));
  printf("sizeof RANKLIST %%zu\n", sizeof(RANKLIST));
  printf("sizeof THRESH %%zu\n", sizeof(THRESH));
  printf("sizeof HIT %%zu\n", sizeof(HIT));
  printf("sizeof HITLIST %%zu\n", sizeof(HITLIST));
%s
  return 0;
}
'''


def _run(includes, incdirs):
    body = "".join(f'  printf("mi.{f} %zu\\n", offsetof(struct mutual_s, {f}));\n' for f in FIELDS_MI)
    body += "".join(f'  printf("data.{f} %zu\\n", offsetof(struct data_s, {f}));\n' for f in FIELDS_DATA)
    body += "".join(f'  printf("hit.{f} %zu\\n", offsetof(HIT, {f}));\n' for f in FIELDS_HIT)
    body += "".join(f'  printf("hitlist.{f} %zu\\n", offsetof(HITLIST, {f}));\n' for f in FIELDS_HITLIST)
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write(PROG % (includes, body))
        exe = os.path.join(td, "p")
        subprocess.run(["gcc", "-w", "-o", exe, src] + [f"-I{d}" for d in incdirs], check=True)
        return subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")
def test_struct_layouts_match_the_reference_header():
    shim = os.path.join(ROOT, "include", "easel_compat")
    mine = _run('#include "rscape_compat.h"', [os.path.join(ROOT, "include"), shim])
    ref = _run('#include "rscape_config.h"\n#include "easel.h"\n#include "correlators.h"',
               [shim, os.path.join(REF, "src"), os.path.join(REF, "lib", "R-view", "src")])
    assert mine == ref, "\n".join(f"{a}   |   {b}" for a, b in zip(mine.splitlines(), ref.splitlines()) if a != b)
