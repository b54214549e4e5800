#!/bin/bash
# Register / stack / shared-memory usage and SASS size of the fused kernels of the in-tree library
# (no GPU needed).  usage: bash profiles/resource_usage.sh [lib.so] > profiles/TAG_resource_usage.txt
LIB=${1:-arboris-python_b200/arboris_b200/lib/libarboris_b200.so}
echo "# cuobjdump --dump-resource-usage $LIB (sm_100a)"
cuobjdump --dump-resource-usage "$LIB" 2>/dev/null | grep -A1 "Function _Z.*k_fused\|Function _Z.*k_state\|Function _Z.*k_update\|Function _Z.*k_integrate" \
  | grep -v "^--" | sed 's/^ *//' | paste - - | sed 's/Function \(_Z[0-9]*\)\([a-z_]*\)[^:]*:/\2:/'
echo
echo "# SASS size per kernel (cuobjdump -sass: instructions x 16 bytes)"
cuobjdump -sass "$LIB" 2>/dev/null | python3 -c '
import re, sys, collections
n = collections.Counter(); name = None
for ln in sys.stdin:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1); continue
    if name and re.match(r"\s*/\*[0-9a-f]+\*/\s+[A-Z@]", ln):
        n[name] += 1
for k in sorted(n):
    if "k_fused" in k:
        print("%-60s %7d instructions  %6.1f KB" % (k, n[k], n[k]*16/1024))
'
