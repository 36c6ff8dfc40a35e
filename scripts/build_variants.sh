#!/bin/bash
# Builds tuning variants of the CUDA library (development aid): scripts/build_variants.sh "name:-DFLAG=1 ..." ...
# Each variant gets its own directory with a copy of the host library beside it (see backend.lib_paths).
set -e
cd "$(dirname "$0")/../abeille_b200/csrc"
rm -rf ../lib/variants
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  mkdir -p ../lib/variants/$name
  make -B EXTRA="$flags" OUT=../lib/variants/$name/libabeille_b200.so >/dev/null 2>&1
  cp ../lib/libabeille_host.so ../lib/variants/$name/
  echo "$name: $(grep -A2 'history_kernelILi1ELb0' ../lib/ptxas.log | grep -E 'registers|spill' | tr '\n' ' ')"
done
make -B >/dev/null 2>&1
