#!/bin/bash
# opcode histogram of one kernel: sass_hist.sh <lib.so> <mangled-name-substring>
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{p=index($0,pat)>0} p' | grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" | awk '{print $NF}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn
