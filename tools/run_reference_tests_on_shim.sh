#!/usr/bin/env bash
# Fidelity check of baseline/ray_shim: the UNMODIFIED reference's own test-suite on top of it.
#   tools/run_reference_tests_on_shim.sh [pytest args]      (CPU only; ~3 min)
# The tests are copied to a scratch directory (the reference tree is read-only) and each file runs
# in its own pytest process, like the reference's test.sh does.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REFERENCE_DIR:-/root/reference}
WORK=$(mktemp -d /tmp/reftests.XXXX)
cp -r "$REF/fed/tests" "$WORK/tests"
cp -r "$REF/tool" "$WORK/tool" 2>/dev/null || true
export PYTHONPATH="$ROOT/baseline/ray_shim:$ROOT/baseline/_ref:${PYTHONPATH:-}"
python "$WORK/tool/generate_tls_certs.py" >/dev/null 2>&1 || true
pass=0; fail=0; failed=()
cd "$WORK/tests"
for f in $(find . -name "test_*.py" | sort); do
  case "$f" in ./client_mode_tests/*) continue;; esac   # Ray-client mode: skipped upstream for ray > 2.4 too
  case "$f" in ./test_utils.py) continue;; esac           # a helper module, no tests inside
  # own session, so that party processes a failing test leaves behind can be reaped as a group
  setsid timeout -k 3 180 python -m pytest -x -q -p no:cacheprovider "$f" "$@" >"$WORK/$(basename "$f").log" 2>&1 &
  pid=$!
  wait $pid; rc=$?
  kill -9 -- -$pid 2>/dev/null
  if [ $rc -eq 0 ]; then
    pass=$((pass+1)); echo "PASS $f"
  else
    fail=$((fail+1)); failed+=("$f"); echo "FAIL $f  (log: $WORK/$(basename "$f").log)"
  fi
done
echo "reference test files on the ray shim: $pass passed, $fail failed ${failed[*]:-}"
# Known and inherent to an in-process stand-in: test_transport_proxy.py::test_send_grpc_with_meta (2nd test of the
# file) relies on every Ray actor being a fresh PROCESS — the reference caches its job config in a module global
# (fed/config.py:46-75) and the test expects the new proxy actor to re-read it from the KV.
unexpected=0
for f in "${failed[@]:-}"; do [ -n "$f" ] && [ "$f" != "./test_transport_proxy.py" ] && unexpected=$((unexpected+1)); done
[ "$unexpected" -eq 0 ]
