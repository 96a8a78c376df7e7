#!/usr/bin/env bash
# Build the C++ party runtime (_fedcore) with a sanitizer and run the runtime / transport /
# multi-party tests against it ("TSAN build of the C++ core").
#   tools/sanitize_core.sh thread     # ThreadSanitizer (default)
#   tools/sanitize_core.sh address    # AddressSanitizer + UBSan
# The interpreter itself is not instrumented, so reports concern our C++ only.
set -euo pipefail
cd "$(dirname "$0")/.."
KIND="${1:-thread}"
OUT="rayfed_b200/build/sanitize-$KIND"
mkdir -p "$OUT"
SUFFIX=$(python -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
INC=$(python -c "import pybind11, sysconfig; print('-I'+pybind11.get_include(), '-I'+sysconfig.get_paths()['include'])")
case "$KIND" in
  thread)  FLAGS="-fsanitize=thread"; LIB=$(g++ -print-file-name=libtsan.so) ;;
  address) FLAGS="-fsanitize=address,undefined -fno-sanitize=vptr"; LIB="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" ;;
  *) echo "unknown sanitizer $KIND" >&2; exit 2 ;;
esac
g++ -O1 -g -fno-omit-frame-pointer $FLAGS -std=c++17 -shared -fPIC -fvisibility=hidden -pthread $INC \
    -Irayfed_b200/csrc/core rayfed_b200/csrc/core/*.cpp -o "$OUT/_fedcore$SUFFIX"
export RAYFED_B200_CORE_SO="$PWD/$OUT/_fedcore$SUFFIX"
export LD_PRELOAD="$(echo $LIB | tr ' ' ':')"
export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 second_deadlock_stack=1 exitcode=0 log_path=$PWD/$OUT/tsan suppressions=$PWD/tools/tsan.supp"
export ASAN_OPTIONS="detect_leaks=0 halt_on_error=0 log_path=$PWD/$OUT/asan"
export UBSAN_OPTIONS="print_stacktrace=1 log_path=$PWD/$OUT/ubsan"
rm -f "$OUT"/tsan.* "$OUT"/asan.* "$OUT"/ubsan.*
TESTS="${SANITIZE_TESTS:-tests/test_core_runtime.py tests/test_transport.py tests/test_two_party.py tests/test_stress.py}"
python -m pytest $TESTS -q -x -m "not gpu" -p no:cacheprovider ${SANITIZE_PYTEST_ARGS:-}
# Reports whose racing stacks never enter _fedcore come from uninstrumented third-party
# code (CPython, grpc's cygrpc, torch) whose own synchronisation TSAN cannot see.
python - "$OUT" "$KIND" <<'PY'
import glob, re, sys
out, kind = sys.argv[1], sys.argv[2]
total = ours = 0
mine = []
for f in glob.glob(f"{out}/tsan.*") + glob.glob(f"{out}/asan.*") + glob.glob(f"{out}/ubsan.*"):
    txt = open(f, errors="replace").read()
    for rep in re.split(r"={18,}\n", txt):
        if "WARNING: ThreadSanitizer" in rep:
            total += 1
            if "_fedcore" in "\n\n".join(rep.split("\n\n")[:2]):
                ours += 1
                mine.append(rep)
        elif "ERROR: AddressSanitizer" in rep or "runtime error" in rep:
            total += 1
            if "_fedcore" in rep or "csrc/core" in rep:
                ours += 1
                mine.append(rep)
print(f"sanitizer={kind}: {total} report(s) in total, {ours} involving _fedcore")
for r in mine[:5]:
    print(r[:3000])
sys.exit(1 if ours else 0)
PY
