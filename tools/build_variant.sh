#!/bin/bash
# usage: build_variant.sh <name> <-D flags...>   -> variants/lib_<name>.so (select it with GIBBS_B200_LIB=...)
name=$1; shift
mkdir -p variants
python - "$name" "$@" <<'PY'
import sys
from gibbssampling_b200 import _build
print(_build.build(out=f"variants/lib_{sys.argv[1]}.so", extra_flags=sys.argv[2:]))
PY
