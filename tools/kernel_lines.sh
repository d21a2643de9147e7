#!/bin/bash
# usage: kernel_lines.sh <object .o> <mangled kernel name> <out.txt>
# nvdisasm -g listing (with //## File/line markers) of ONE kernel, the input prof_by_line.py / prof_by_region.py expect
obj=$1; name=$2; out=$3
tmp=$(mktemp -d)
(cd $tmp && cuobjdump -xelf all $obj > /dev/null 2>&1)
cub=$(ls $tmp/*.cubin | head -1)
nvdisasm -g $cub 2>/dev/null | awk -v n="$name" '
  /^\/\/--------------------- \.text\./ { on = (index($0, ".text." n " ") > 0 || $0 ~ ("\\.text\\." n "[ \t]*-*$")) }
  on { print }' > $out
rm -rf $tmp
wc -l $out
