#!/bin/bash
# A/B timing of analysis builds: tools/ab.sh name1 name2 ... (mpmavatar_b200/libmpm_b200_<name>.so; "base" = the product build)
for n in "$@"; do
  lib=mpmavatar_b200/libmpm_b200_$n.so; [ "$n" = base ] && lib=mpmavatar_b200/libmpm_b200.so
  echo "== $n"; MPM_B200_LIB=$PWD/$lib python tools/quick_time.py ${SCENE:-c3} 400 2>&1 | grep -E "substeps/s|us/substep|finite"
done
