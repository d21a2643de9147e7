#!/bin/bash
# usage: build_variant.sh <name> [--only unit1,unit2] <-D flags...>   -> variants/lib_<name>.so (select it with GIBBS_B200_LIB=...)
# --only: the units (object names of _build.py, e.g. launch_chain_t4,launch_init_smem) the flags apply to; the rest is
# linked from the in-tree build
name=$1; shift
mkdir -p variants
python - "$name" "$@" <<'PY'
import sys
from gibbssampling_b200 import _build
args = sys.argv[2:]
only = None
if args and args[0] == "--only":
    only = args[1].split(","); args = args[2:]
print(_build.build(out=f"variants/lib_{sys.argv[1]}.so", extra_flags=args, only=only))
PY
