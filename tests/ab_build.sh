#!/bin/bash
# Build libcoddvec.so of another git ref (a WIP branch, an older commit) beside the working tree's own library, for
# same-box A/B runs on the GPU:   tests/ab_build.sh wip/k2-smem-queries   ->  ab_libs/wip_k2-smem-queries.so
# On the box:  CODDVEC_LIB=$PWD/ab_libs/<name>.so python tests/quick_probe.py   (native.library_path honours it).
# ab_libs/*.so is git-ignored (*.so) but travels with gpurun like the in-tree library.
set -euo pipefail
ref=${1:?usage: tests/ab_build.sh <git-ref>}
root=$(cd "$(dirname "$0")/.." && pwd)
name=$(echo "$ref" | tr '/' '_')
mkdir -p "$root/build/wt"
tmp=$(mktemp -d "$root/build/wt/ab_tree.XXXXXX")
git -C "$root" worktree add -f --detach "$tmp" "$ref" > /dev/null
( cd "$tmp" && python -m codd_query_engine_b200.build > /dev/null )
mkdir -p "$root/ab_libs"
cp "$tmp/codd_query_engine_b200/libcoddvec.so" "$root/ab_libs/$name.so"
git -C "$root" worktree remove --force "$tmp"
git -C "$root" worktree prune
echo "built $root/ab_libs/$name.so from $ref"
