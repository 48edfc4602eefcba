#!/bin/sh
# build everything; exit non-zero if the CUDA library did not compile
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { grep -E "error" /tmp/build.log | head; echo BUILD FAILED; exit 1; }
grep -E "error" /tmp/build.log | head
echo BUILD OK
