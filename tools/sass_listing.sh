#!/bin/bash
# tools/sass_listing.sh — per hot kernel of libsoundkit_cuda.so: counts of the SASS mnemonics that show what the code is made of
# (FFMA2 packed FMAs, UBLKCP bulk async copies, SYNCS mbarrier ops, LDGSTS cp.async, LDS / STS widths).  Output: profiles/*_sass_mnemonics.txt
LIB=${1:-soundkit_b200/libsoundkit_cuda.so}
cuobjdump -sass "$LIB" | awk '
/Function : /{ if (name != "") flush(); name=$3; delete c; next }
/^[ \t]+\/\*[0-9a-f]+\*\//{ op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); split(op, parts, "."); base=parts[1];
   if (op ~ /^FFMA2/) c["FFMA2"]++; else if (op ~ /^FFMA/) c["FFMA"]++;
   if (op ~ /^UBLKCP/) c["UBLKCP"]++; if (op ~ /^SYNCS/) c["SYNCS"]++; if (op ~ /^LDGSTS/) c["LDGSTS"]++;
   if (op ~ /^LDS\.128/) c["LDS.128"]++; else if (op ~ /^LDS/) c["LDS"]++;
   if (op ~ /^STS\.128/) c["STS.128"]++; else if (op ~ /^STS/) c["STS"]++;
   if (op ~ /^LDG.*128/) c["LDG.128"]++; if (op ~ /^STG.*128/) c["STG.128"]++; if (op ~ /^SHFL/) c["SHFL"]++; if (op ~ /^LDCU/) c["LDCU"]++;
   if (op ~ /^PRMT/) c["PRMT"]++; c["total"]++ }
function flush(   k, line) { line=""; n=split("total FFMA2 FFMA LDCU LDS.128 LDS STS.128 STS LDG.128 STG.128 LDGSTS UBLKCP SYNCS SHFL PRMT", keys, " ");
   for (k=1;k<=n;k++) if (c[keys[k]] > 0) line = line sprintf(" %s=%d", keys[k], c[keys[k]]); print name ":" line }
END { flush() }' | c++filt | grep -E "resample_|g711_|flat_kernel|deint_|interleave_kernel|norm48k_tile|g726_|g722_|mixdown|quantize_mono" | sort
