#!/bin/bash
# Time the UNMODIFIED reference (nedtaylor/athena, Fortran) on the host CPU for the bench.py
# workload.
#
# The development image has no Fortran compiler and the reference's three fpm dependencies
# (coreutils v0.1.0, diffstruc v1.2.0, graphstruc v0.2.1; fpm.toml:19-21) are fetched from
# GitHub at build time, so this cannot run there; bench.py --impl reference therefore times the
# line-by-line C restatement (oracle/).  On a machine with gfortran >= 14, fpm and network
# access this script builds athena with the release profile (-O3 -march=native,
# CMakeLists.txt:174 / fpm --profile release), drops baseline/bench_msgpass.f90 into the
# tree as an fpm app and runs it on ONE core (the reference has no working threading:
# CMakeLists.txt:38-42, athena_network_sub.f90:27-31).
#
#   ATHENA_REF=/path/to/athena  baseline/run_reference.sh [graphs] [epochs]
#
# Prints one JSON line: {"impl": "reference-fortran", "value": edges/s, ...}.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${ATHENA_REF:-/root/reference}"
GRAPHS="${1:-4096}"
EPOCHS="${2:-3}"
for tool in gfortran fpm; do
  if ! command -v "$tool" >/dev/null 2>&1; then
    echo "{\"impl\": \"reference-fortran\", \"unavailable\": \"$tool not found on this host\"}"
    exit 0
  fi
done
WORK="$(mktemp -d)"
trap 'rm -rf "$WORK"' EXIT
cp -r "$REF" "$WORK/athena"          # the reference tree is read-only: build in a copy
mkdir -p "$WORK/athena/app"
cp "$HERE/bench_msgpass.f90" "$WORK/athena/app/bench_msgpass.f90"
cd "$WORK/athena"
export OMP_NUM_THREADS=1
fpm build --profile release --flag "-march=native" >"$WORK/build.log" 2>&1 || {
  echo "{\"impl\": \"reference-fortran\", \"unavailable\": \"fpm build failed (see build log; the dependencies need network access)\"}"
  tail -20 "$WORK/build.log" >&2
  exit 0
}
taskset -c 0 fpm run --profile release --flag "-march=native" bench_msgpass -- "$GRAPHS" "$EPOCHS"
