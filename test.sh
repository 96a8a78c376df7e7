#!/usr/bin/env bash
# Test runner (parity: the reference's test.sh — there: generate certs, then one pytest
# process per file).  Here the build step comes first; each test file still gets a fresh
# interpreter so that a wedged party process cannot poison the next file.
#   ./test.sh            CPU suite
#   ./test.sh gpu        GPU suite (B200)
#   ./test.sh sanitize   TSAN + ASAN runs of the C++ core
set -euo pipefail
cd "$(dirname "$0")"
MODE="${1:-cpu}"
python __graft_entry__.py >/dev/null
case "$MODE" in
  cpu)      MARK="not gpu" ;;
  gpu)      MARK="gpu" ;;
  sanitize) bash tools/sanitize_core.sh thread && SANITIZE_PYTEST_ARGS="-k not(grpc)" bash tools/sanitize_core.sh address; exit $? ;;
  *) echo "usage: $0 [cpu|gpu|sanitize]" >&2; exit 2 ;;
esac
status=0
for f in tests/test_*.py; do
  echo "== $f"
  python -m pytest "$f" -q -m "$MARK" -p no:cacheprovider || { rc=$?; [ "$rc" = 5 ] || status=$rc; }
done
exit $status
