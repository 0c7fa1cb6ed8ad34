#!/bin/bash
# tools/sass_of.sh PATTERN [lib] -> SASS of the first kernel whose mangled name contains PATTERN
LIB=${2:-iactrace_b200/csrc/libiactrace_b200.so}
NAME=$(cuobjdump -elf "$LIB" 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*$1[A-Za-z0-9_]*" | grep -v _param | sort -u | head -1)
echo "// $NAME"
cuobjdump -sass -fun "$NAME" "$LIB" 2>/dev/null | grep -E "^\s*/\*[0-9a-f]{4,6}\*/"
