#!/bin/bash
# usage: benchmarks/sass_mix.sh <object-or-so> <kernel-name-regex>   -- instruction mix of the first matching kernel
f=$1; pat=$2
name=$(cuobjdump -sass "$f" | grep -E "Function : .*${pat}" | head -1 | sed 's/.*Function : //')
echo "kernel: $name"
cuobjdump -sass -fun "$name" "$f" | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+).*/\2/' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -25
