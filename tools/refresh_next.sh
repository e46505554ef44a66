#!/bin/bash
# Regenerate next/r2-prep.diff (= git diff main r2-prep, without next/ itself) on the MAIN tree, from any cwd.
set -e
cd "$(git -C "$(dirname "$0")/.." rev-parse --show-toplevel)"
[ "$(git rev-parse --abbrev-ref HEAD)" = "main" ] || { echo "run from the main checkout"; exit 1; }
git diff main r2-prep -- . ':!next' > next/r2-prep.diff
git apply --check next/r2-prep.diff
git add next && (git commit -qm "next/: refresh the r2-prep patch" || true)
git status --short | head -3
