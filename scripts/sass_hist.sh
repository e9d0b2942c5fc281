#!/bin/bash
# usage: sass_hist.sh <lib.so> <mangled-function>   -> static opcode histogram of one kernel
cuobjdump -sass -fun "$2" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | tr '\n' ' '; echo
