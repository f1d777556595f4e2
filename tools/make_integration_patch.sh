#!/bin/bash
# Regenerates integration/sfsim.patch from the shims in integration/clj and the small edits of the reference
# (build.clj, deps.edn, Makefile, scripts/packr-config-linux.json).  Needs the reference checkout (default
# /root/reference).  tests/test_integration_patch.py checks that the committed patch still applies.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
for f in build.clj deps.edn Makefile scripts/packr-config-linux.json; do
  mkdir -p $W/a/$(dirname $f) $W/b/$(dirname $f)
  cp $REF/$f $W/a/$f
  cp $REF/$f $W/b/$f
done
(cd $W/b && git init -q . && git apply $HERE/integration/sfsim.patch && rm -rf .git)
cp $HERE/integration/clj/sfsim/atmosphere_cuda.clj $HERE/integration/clj/sfsim/globe_cuda.clj $W/b/src/clj/sfsim/   # the shims are edited there
(cd $W && git diff --no-index --no-color a b || true) | sed 's@^diff --git a/a/@diff --git a/@; s@^diff --git a/b/@diff --git a/@; s@ b/b/@ b/@; s@^--- a/a/@--- a/@; s@^+++ b/b/@+++ b/@'
rm -rf $W
