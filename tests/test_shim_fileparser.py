"""The esl_fileparser subset of the Easel shim (r-scape_b200/host/easel_shim.c): what the reference's readers of token files use
(cov_ReadNullHistogram, src/covariation.c:1739-1760: Open, SetCommentChar, NextLine, GetTokenOnLine, Close)."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ESL_OK, ESL_EOL, ESL_EOF, ESL_ENOTFOUND = 0, 2, 3, 6


def _lib():
    lib = C.CDLL(os.path.join(ROOT, "r-scape_b200", "librscape_b200_host.so"))
    lib.esl_fileparser_Open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.esl_fileparser_SetCommentChar.argtypes = [C.c_void_p, C.c_char]
    lib.esl_fileparser_NextLine.argtypes = [C.c_void_p]
    lib.esl_fileparser_GetTokenOnLine.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
    lib.esl_fileparser_Close.argtypes = [C.c_void_p]
    lib.esl_fileparser_Close.restype = None
    return lib


def _parse(lib, path, comment=b"#"):
    efp = C.c_void_p()
    assert lib.esl_fileparser_Open(path.encode(), None, C.byref(efp)) == ESL_OK
    lib.esl_fileparser_SetCommentChar(efp, comment)
    rows = []
    while lib.esl_fileparser_NextLine(efp) == ESL_OK:
        row, tok, n = [], C.c_char_p(), C.c_int()
        while lib.esl_fileparser_GetTokenOnLine(efp, C.byref(tok), C.byref(n)) == ESL_OK:
            assert n.value == len(tok.value)
            row.append(tok.value.decode())
        assert lib.esl_fileparser_GetTokenOnLine(efp, C.byref(tok), None) == ESL_EOL      # stays at the end of the line
        rows.append(row)
    assert lib.esl_fileparser_NextLine(efp) == ESL_EOF
    lib.esl_fileparser_Close(efp)
    return rows


def test_tokens_comments_blank_lines_and_long_lines(tmp_path):
    lib = _lib()
    long_tok = "x" * 1000
    p = tmp_path / "f.txt"
    p.write_bytes(("# header\n\n  -10.000000 0\n-9.950000\t17   # trailing comment\r\n   \t \n#only a comment\n" + long_tok + " 3 4\nlast 1").encode())
    assert _parse(lib, str(p)) == [["-10.000000", "0"], ["-9.950000", "17"], [long_tok, "3", "4"], ["last", "1"]]


def test_empty_file_and_missing_file(tmp_path):
    lib = _lib()
    p = tmp_path / "empty.txt"
    p.write_text("")
    assert _parse(lib, str(p)) == []
    efp = C.c_void_p()
    assert lib.esl_fileparser_Open(str(tmp_path / "nope.txt").encode(), None, C.byref(efp)) == ESL_ENOTFOUND
    assert not efp.value
