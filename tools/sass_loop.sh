#!/bin/bash
# prints the SASS of one forest-kernel instantiation (default: packed kind 3, EI path)   usage: tools/sass_loop.sh [lib] [mangled-substring]
lib=${1:-hypermapper_b200/csrc/libhm_b200.so}; fn=${2:-hm_forest_kernelILb0ELb0ELi1024ELi3E}
cuobjdump -sass $lib 2>/dev/null | awk -v fn="$fn" '/Function :/{p=index($0,fn)>0} p' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
