#!/bin/bash
# Builds the engine of the last commit as build/libpvv_prev.so (A/B partner of the working tree: tools/ab_perf.sh).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
rm -rf /tmp/pvv_prev && git -C "$ROOT" worktree prune && git -C "$ROOT" worktree add -f /tmp/pvv_prev HEAD > /dev/null 2>&1
(cd /tmp/pvv_prev/py_vollib_vectorized_b200 && python build.py --force > /dev/null)
mkdir -p "$ROOT/py_vollib_vectorized_b200/build"
cp /tmp/pvv_prev/py_vollib_vectorized_b200/libpvv.so "$ROOT/py_vollib_vectorized_b200/build/libpvv_prev.so"
git -C "$ROOT" worktree remove --force /tmp/pvv_prev
echo "build/libpvv_prev.so = $(git -C "$ROOT" rev-parse --short HEAD)"
