#!/bin/bash
# usage: scripts/sass_of.sh <lib.so> <cubin-name-substring> <kernel-name-substring>  -> SASS of one kernel on stdout
set -e
D=$(mktemp -d); cd $D
cuobjdump -xelf all "$1" >/dev/null
C=$(ls *.cubin | grep "$2" | head -1)
cuobjdump -sass $C | awk -v k="$3" '/Function : /{f=(index($0,k)>0)} f'
