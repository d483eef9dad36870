#!/bin/sh
# Copies the engine sources next to the glue so that `R CMD INSTALL pkg` (or `R CMD build pkg`) finds them in pkg/src.
#   sh pkg/bootstrap.sh && R CMD INSTALL pkg
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(dirname "$here")
cp "$root"/spnngp_b200/csrc/*.cu "$root"/spnngp_b200/csrc/*.cuh "$root"/spnngp_b200/csrc/*.h "$here"/src/
cp "$root"/spnngp_b200/rglue/spnngp_glue.cpp "$here"/src/
cp "$root"/include/nngp_b200.h "$here"/src/
# the sources include the public header by its repository-relative path
sed -i 's#"../../include/nngp_b200.h"#"nngp_b200.h"#' "$here"/src/*.h "$here"/src/*.cuh "$here"/src/*.cu "$here"/src/*.cpp 2>/dev/null || true
echo "pkg/src is ready: $(ls "$here"/src | wc -l) files"
