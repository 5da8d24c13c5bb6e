#!/bin/bash
# SASS identity check: do the kernels of one csrc file compile to the same machine code in the working tree as in a git
# revision (e.g. the last revision whose tests ran on a B200)?  Lets diagnostics / staged template variants be added without
# re-validating the default instantiations: identical SASS = identical behaviour.
#   scripts/sass_diff.sh <git-rev> <file.cu>       (kernels present on only one side are listed, not compared)
set -euo pipefail
rev=$1; f=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/old/stcat_b200/csrc" "$tmp/old/include"
for h in $(git -C "$root" ls-tree --name-only "$rev" stcat_b200/csrc/ | grep -E '\.cuh$'); do git -C "$root" show "$rev:$h" > "$tmp/old/$h"; done
git -C "$root" show "$rev:include/stcat_b200.h" > "$tmp/old/include/stcat_b200.h"
git -C "$root" show "$rev:stcat_b200/csrc/$f" > "$tmp/old/stcat_b200/csrc/$f"
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
(cd "$tmp/old/stcat_b200/csrc" && nvcc $flags -c "$f" -o "$tmp/old.o")
(cd "$root/stcat_b200/csrc" && nvcc $flags -c "$f" -o "$tmp/new.o")
funs() { cuobjdump -sass "$1" | grep -oE 'Function : \S+' | awk '{print $3}' | sort; }
body() { cuobjdump -sass -fun "$2" "$1" | grep -E '^\s+/\*[0-9a-f]{4,8}\*/' | sed -E 's#/\*[0-9a-f]+\*/##g; s/^\s+//'; }
# template-argument lists may have grown (new trailing bool parameters defaulting to the old behaviour): match by demangled
# prefix when the exact symbol is missing
rc=0
for fn in $(funs "$tmp/old.o"); do
    cand=$fn
    if ! funs "$tmp/new.o" | grep -qx "$fn"; then
        pre=$(echo "$fn" | sed -E 's/EEEv.*$//')
        cand=$(funs "$tmp/new.o" | grep -E "^${pre}E(Lb0E)*EEv" | head -1 || true)
    fi
    if [ -z "$cand" ]; then echo "MISSING  $(echo "$fn" | c++filt | cut -c1-100)"; rc=1; continue; fi
    if diff -q <(body "$tmp/old.o" "$fn") <(body "$tmp/new.o" "$cand") > /dev/null; then
        echo "same     $(echo "$cand" | c++filt | cut -c1-100)"
    else
        echo "DIFFERS  $(echo "$cand" | c++filt | cut -c1-100)"; rc=1
    fi
done
rm -rf "$tmp"
exit $rc
