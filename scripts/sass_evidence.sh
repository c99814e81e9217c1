#!/bin/bash
# Count the Blackwell-specific SASS mnemonics per kernel of the shipped library (tcgen05.mma -> UTCHMMA,
# tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, tcgen05.commit -> UTCBAR) and show one excerpt per kernel.
SO=${1:-xmask3d_b200/libxm3d.so}
OUT=${2:-profiles/r02_sass_evidence.txt}
{
echo "# cuobjdump -sass $SO  (sm_100a) — count, kernel, mnemonic"
cuobjdump -sass "$SO" 2>/dev/null | awk '/Function :/{fn=$3} /UTC[A-Z]*MMA|UTMALDG|UTMASTG|LDTM|STTM|UBLKCP|UTCBAR/{
  if (match($0,/UTC[A-Z]*MMA[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|STTM[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*/)) {c[fn" "substr($0,RSTART,RLENGTH)]++; if (!(fn in ex)) ex[fn]=$0}}
  END{for (x in c) print c[x], x; print "# first matching instruction per kernel"; for (f in ex) print f": "ex[f]}' | sort -k2
} > "$OUT"
echo "wrote $OUT"
