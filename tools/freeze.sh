#!/bin/bash
# Freeze the working tree into .frozen/<tag>/ so that a queued gpurun job runs exactly this state while
# editing goes on (gpurun snapshots /root/repo when the box arrives, not when the call is queued).
# usage: tools/freeze.sh <tag>   ->  job command:  cd .frozen/<tag> && bash tools/<script>.sh
set -e
tag=$1; root=$(cd "$(dirname "$0")/.." && pwd)
rm -rf "$root/.frozen/$tag"; mkdir -p "$root/.frozen/$tag"
cd "$root"
tar --exclude=./.git --exclude=./gpurun_out --exclude=./.frozen --exclude='__pycache__' --exclude=./.pytest_cache -cf - . | tar -xf - -C "$root/.frozen/$tag"
ln -s ../../gpurun_out "$root/.frozen/$tag/gpurun_out"
echo "frozen -> .frozen/$tag"
