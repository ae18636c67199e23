#!/bin/bash
# Tensor-core / async-copy instructions per kernel of the built library (B200_PROFILING.md: the SASS mnemonics that
# prove DMMA, LDGSTS, UBLKCP (cp.async.bulk, TMA unit) and mbarrier use).  Runs without a GPU.
so=${1:-axctdprocessor_b200/libaxctd.so}
cuobjdump -sass "$so" 2>/dev/null | awk '
/Function : /{fn=$3}
{ for (i = 1; i <= NF; i++) if ($i ~ /^(UBLKCP|UTMALDG|SYNCS|LDGSTS|DMMA|IMMA|IDP\.2A|DFMA|FFMA|FADD|I2F|F2F)/) { m=$i; sub(/\..*/, "", m); if ($i ~ /^IDP/) m="IDP.2A"; c[fn" "m]++ } }
END { for (k in c) print k, c[k] }' | sort | awk '{ if ($1 != last) { if (last != "") print line; line = $1 ":"; last = $1 } line = line " " $2 "=" $3 } END { print line }' | c++filt | grep -E "k_demod_fused|k_stats_tones|k_tone_windows|k_demod_ws|k_decim_fused"
