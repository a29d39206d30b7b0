#!/usr/bin/env bash
# usage: tools/sass.sh <kernel-name-substring> [lib]   -> SASS of the first matching kernel
LIB="${2:-$(dirname "$0")/../compressed-fm-index-implementation-with-learned-optimizations_b200/libcsfm.so}"
cuobjdump -sass "$LIB" | awk -v pat="$1" '
  /Function :/ { on = (index($0, pat) > 0) }
  on { print }'
