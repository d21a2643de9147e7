#!/bin/bash
# usage: sass_excerpt.sh <object .o> <mangled kernel name> <mnemonic regex> <out.txt>
# The lines of ONE kernel's SASS (cuobjdump -sass) that match the regex, with their instruction addresses, plus a count per
# mnemonic: the evidence that a kernel uses TMA bulk copies / mbarriers / cluster barriers / warp reductions.
obj=$1; name=$2; re=$3; out=$4
cuobjdump -sass $obj | awk -v n="$name" '
  /Function :/ { on = (index($0, n) > 0) }
  on && /^[ \t]*\/\*[0-9a-f]+\*\// { print }' > /tmp/sass_one.$$
{
  echo "# $name ($(basename $obj)), $(wc -l < /tmp/sass_one.$$) SASS instructions; lines matching /$re/"
  echo "# counts per mnemonic:"
  grep -E "$re" /tmp/sass_one.$$ | sed -E 's/^[ \t]*\/\*[0-9a-f]+\*\/[ \t]+(@!?U?P[0-9T]+ )?//' | awk '{print $1}' | sort | uniq -c | sort -rn | sed 's/^/#   /'
  grep -E "$re" /tmp/sass_one.$$ | sed -E 's/[ \t]+\/\* 0x[0-9a-f]+ \*\/$//'
} > $out
rm -f /tmp/sass_one.$$
wc -l $out
