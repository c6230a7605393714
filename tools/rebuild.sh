#!/bin/bash
# rebuild the libraries (extra env, e.g. SVB_BUILD_KO=1, passes through) and print the ptxas summary of kernels matching $1
cd /root/repo
python -c "
import sys; sys.path.insert(0,'.')
import importlib
b = importlib.import_module('iuvl_b200.build')
b.build(force=('${FORCE:-0}'=='1'))" 2>&1 | grep -v "^ptxas\|^    " | tail -15
if [ -n "$1" ]; then grep -A3 "$1" interactable-unified-vision-language_b200/build/attention_tc.ptxas.log | grep -E "Compiling|spill|Used" | sed 's/_ZN3svb[0-9a-zA-Z_]*attn/attn/' | cut -c1-150; fi
