#!/bin/bash
# Where does the record epilogue's time go?  Variants of the library with parts of it switched off (timing only).
cd "$(dirname "$0")/.."
A=r-scape_b200/build/alt
python tools/gram_time.py ssu 1,2,4
for v in rowform noepi nomma nomma_row; do RSCAPE_B200_LIB=$PWD/$A/$v.so python tools/gram_time.py ssu 1,2,4; done
