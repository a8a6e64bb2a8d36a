#!/bin/bash
# Static SASS statistics of the block-kernel instantiations: total instructions and the mix of one kernel.
#   tools/sass_stats.sh [lib.so] [substring of the mangled kernel name]
LIB=${1:-skrample_b200/csrc/libskrample_b200.so}
cuobjdump -sass "$LIB" | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ /{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -n | grep -E "block_kernel|step_kernel" | cut -c1-140
if [ -n "$2" ]; then
  cuobjdump -sass "$LIB" | awk -v pat="$2" '/Function :/{on = index($3, pat) > 0} on && /^ +\/\*[0-9a-f]+\*\/ /{print}' \
    | grep -o "^ *\/\*[0-9a-f]*\*\/ *\(@!\?U\?P[0-9T] \)\?[A-Z0-9_.]*" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -25
fi
