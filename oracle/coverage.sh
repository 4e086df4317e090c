#!/bin/bash
# TEST INFRASTRUCTURE: which lines of the restatement (nuclear-sim_b200/csrc/plant/*.h) do the live-reference fixtures execute?
# Builds the host restatement with gcov instrumentation in a scratch directory, runs the fixture tests against it and
# prints per-header line coverage plus every line that never ran.  The regular oracle library is put back afterwards.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
COV="${1:-/tmp/nps_cov}"
mkdir -p "$COV"; rm -f "$COV"/*.gcda "$COV"/*.gcov
make -C "$ROOT/oracle" > /dev/null
g++ -O0 --coverage -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -fno-builtin-pow -I"$ROOT/nuclear-sim_b200/csrc/plant" \
    -c "$ROOT/oracle/cpu_port.cpp" -o "$COV/cpu_port.o"
g++ --coverage -shared -o "$COV/libnps_oracle_cov.so" "$COV/cpu_port.o" -lm
cp "$ROOT/oracle/_build/libnps_oracle.so" "$COV/libnps_oracle_plain.so"
trap 'cp "$COV/libnps_oracle_plain.so" "$ROOT/oracle/_build/libnps_oracle.so"' EXIT
cp "$COV/libnps_oracle_cov.so" "$ROOT/oracle/_build/libnps_oracle.so"
(cd "$ROOT" && python -m pytest tests/test_oracle_vs_reference_golden.py tests/test_maintenance_host.py -q | tail -1)
(cd "$COV" && gcov -o . "$ROOT/oracle/cpu_port.cpp" > gcov_summary.txt 2>&1)
tot_un=0; tot_ex=0
for f in "$COV"/*.h.gcov; do
  h=$(basename "$f" .gcov)
  case "$h" in fastpow.h|hd.h|state.h|prefetch.h|rng.h) continue;; esac
  grep -q "csrc/plant/$h" "$f" || continue
  un=$(grep -c '#####' "$f" || true); ex=$(grep -cE '^\s+[0-9]+\*?:' "$f" || true)
  tot_un=$((tot_un + un)); tot_ex=$((tot_ex + ex))
  printf "%-20s %4d of %4d lines never executed\n" "$h" "$un" "$((un + ex))"
  grep '#####' "$f" | cut -c1-150 | sed 's/^/      /' || true
done
echo "total: $tot_un of $((tot_un + tot_ex)) lines never executed by any live-reference fixture"
