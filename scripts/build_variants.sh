#!/bin/bash
# usage: scripts/build_variants.sh name1:"-DA=1 -DB=2" name2:"" ...   -> topoflow_glacier_b200/lib/variants/<name>.so
mkdir -p topoflow_glacier_b200/lib/variants
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  python -m topoflow_glacier_b200.build $defs --out=topoflow_glacier_b200/lib/variants/$name.so > /dev/null && echo "built $name ($defs)" &
done
wait
