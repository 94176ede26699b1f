#!/bin/sh
# apply_glue.sh SRC_DIR — prepares the reference's src/ for the B200 binding.
#
# Removes from SRC_DIR/microclimfCpp.cpp the twelve functions that rglue/microclimf_glue.cpp replaces (each from its
# `// [[Rcpp::export]]` marker to the closing brace in column 0), keeps a backup next to it, and copies the glue file in.
# Signatures are unchanged, so src/RcppExports.cpp and R/RcppExports.R need no edit (Rcpp::compileAttributes() would
# regenerate them identically).  Then add rglue/Makevars to src/ and R CMD INSTALL as usual.
set -e
SRC=${1:?usage: apply_glue.sh SRC_DIR}
HERE=$(cd "$(dirname "$0")" && pwd)
F="$SRC/microclimfCpp.cpp"
[ -f "$F" ] || { echo "$F not found" >&2; exit 1; }
cp "$F" "$F.orig"
awk '
BEGIN {
    n = split("runmicro1Cpp runmicro2Cpp runmicro3Cpp runmicro4Cpp runbioclim1Cpp runbioclim2Cpp runbioclim3Cpp runbioclim4Cpp gridmodelsnow1 gridmodelsnow2 gridmicrosnow1 gridmicrosnow2", names, " ")
    for (i = 1; i <= n; ++i) want["List " names[i] "("] = names[i]
    held = 0; skipping = 0; removed = 0
}
{
    if (skipping) {                       # inside a replaced function: drop up to the closing brace in column 0
        if ($0 ~ /^}/) { skipping = 0; removed++ }
        next
    }
    if ($0 ~ /^\/\/ \[\[Rcpp::export\]\]/) {   # hold the marker until the next line tells whose it is
        if (held) print heldline
        held = 1; heldline = $0
        next
    }
    hit = 0
    for (k in want) if (index($0, k) == 1) hit = 1
    if (hit) { held = 0; skipping = 1; next }   # drop marker + function
    if (held) { print heldline; held = 0 }
    print
}
END {
    if (held) print heldline
    if (removed != 12) { print "apply_glue: removed " removed " of 12 functions" > "/dev/stderr"; exit 2 }
}' "$F.orig" > "$F"
cp "$HERE/microclimf_glue.cpp" "$SRC/microclimf_glue.cpp"
echo "apply_glue: 12 functions of $F replaced by microclimf_glue.cpp (backup: $F.orig)"
