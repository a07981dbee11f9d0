#!/bin/bash
# resource usage and opcode counts of the built library -> profiles/rNN_sass_summary.txt
LIB=${1:-sybil_b200/csrc/libsybilgpu.so}
echo "# $LIB: resource usage (cuobjdump --dump-resource-usage) of the scan / staging kernels"
cuobjdump --dump-resource-usage $LIB 2>/dev/null | grep -A1 -E "Function.*(scan_kernel|stats_kernel|distinct_kernel)" | grep -v "^--" | sed 's/^ *//'
echo
echo "# opcode counts (cuobjdump -sass, whole library: both builds of the kernel unit, 7 scan_kernel instantiations each)"
cuobjdump -sass $LIB 2>/dev/null | grep -oE '^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?[A-Z][A-Z0-9_.]*' | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn \
  | grep -E ' (UTMALDG|UTMAPF|SYNCS|ATOMS|ATOMG|REDG|RED|REDUX|LDS|STS|LDL|STL|LDG|STG|SHFL|VOTE|MATCH|BAR|IMAD|VIADDMNMX|VIMNMX|R2P|PRMT|LOP3|FENCE|MEMBAR|BSSY|BSYNC|BRA|HMMA|IMMA|UTCHMMA|UTCQMMA)$'
