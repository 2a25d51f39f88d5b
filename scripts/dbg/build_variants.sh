#!/bin/bash
# usage: scripts/dbg/build_variants.sh name1="-DFLAG ..." name2="..."   -> weaklysuperviseddl_b200/libx_<name>.so (experiments)
set -e
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  WSDL_NVCC_EXTRA="$flags" python -m weaklysuperviseddl_b200.build --force > /dev/null
  cp weaklysuperviseddl_b200/libwsdl_b200.so "weaklysuperviseddl_b200/libx_$name.so"
  echo "built $name [$flags]"
done
python -m weaklysuperviseddl_b200.build --force > /dev/null
