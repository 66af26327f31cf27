#!/bin/bash
# scripts/sass_of.sh <function-name-substring> : SASS of one kernel of the in-tree library, one instruction per line
cuobjdump -sass /root/repo/genomicsbench_b200/lib/libbsw_b200.so | awk -v pat="$1" '/Function : /{p = index($0, pat) > 0} p' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
