#!/bin/bash
# SASS of one kernel of the built library, instruction lines only.   bash tools/sass_of.sh <mangled-name-fragment>
cuobjdump -sass "$(dirname "$0")/../verifybamid_b200/libvb2llk.so" | awk -v pat="$1" '/Function : /{f = index($0, pat) > 0} f' | grep -v "^\s*/\* 0x" | awk 'NF' | sed 's/^\s*//' | cut -c1-120
