#!/bin/sh
# Type-checks r-package/src/b200als_shim.c against include/b200als.h and stand-in R headers (stub_R/, real signatures,
# declarations only) in an image without R.  It proves the shim and the C ABI agree on every call; it does not run R.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
gcc -std=c11 -Wall -Wextra -Wno-unused-parameter -Wno-cast-function-type -Werror -fsyntax-only -I "$HERE/stub_R" -I "$HERE/../../include" "$HERE/../src/b200als_shim.c"
echo "b200als_shim.c: type-check ok"
