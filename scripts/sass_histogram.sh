#!/bin/bash
# static SASS mnemonic histogram of the production kernels (cuobjdump -sass on the built library)
LIB=bajes_b200/csrc/libbajes_b200.so
for pat in 'fullgrid_tile_kernelILi3ELi0ELi0ELb1ELb0ELb1E' 'fullgrid_window_kernelILi3ELi0ELb0E' 'roq_kernelILi0E' 'relbin_kernelILi3ELi0ELb1E' 'epilogue_kernel' 'prologue_kernel'; do
  fn=$(cuobjdump -sass $LIB | grep -o "Function : .*$pat[^ ]*" | head -1 | sed 's/Function : //')
  echo "== $fn"
  cuobjdump -sass -fun "$fn" $LIB 2>/dev/null | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9]+\s+)?//' | awk '{print $1}' | sed -E 's/;$//' | awk -F. '{k=$1; if ($1=="MUFU"||$1=="DMMA"||$1=="UBLKCP"||$1=="SYNCS"||$1=="LDS"||$1=="LDG"||$1=="F2F"||$1=="I2FP") k=$1"."$2; c[k]++} END {for (k in c) printf "%6d %s\n", c[k], k}' | sort -rn | head -44
done
