#!/usr/bin/env bash
# regenerates profiles/r2_sass_mnemonics.md from the built objects
cd "$(dirname "$0")/.."
for f in kernels_batch bq_imma bq_umma search build; do
  echo "## $f.cu"; cuobjdump -sass jvector_b200/lib/obj/$f.cu.o | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sort | uniq -c | sort -rn | head -40
done
