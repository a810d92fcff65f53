#!/bin/bash
# Race / memory / sync checking of the kernels (the reference has none).
# Run on a GPU box: tools/sanitize.sh [memcheck|racecheck|synccheck] [logfile]
#   racecheck runs the RSDL_RACECHECK build (csrc/build/variants/racecheck): same kernels
#   plus a named barrier on the index -> consumer hand-off, which the tool can model
#   (it does not model mbarrier arrive / try_wait); see profiles/README.md "sanitizers".
set -uo pipefail
cd "$(dirname "$0")/.."
tool="${1:-memcheck}"
log="${2:-/dev/stdout}"
ext=()
if [ "$tool" = "racecheck" ]; then
  python -c "from ray_shuffling_data_loader_b200 import _build; print(_build.build_variant('racecheck', ['RSDL_RACECHECK=1']))"
  ext=(--ext ray_shuffling_data_loader_b200/csrc/build/variants/racecheck/_C$(python -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))"))
fi
rc=0
for args in "--cols 21 --mode 0" "--cols 21 --mode 3" "--cols 21 --mode 4" "--cols 70 --mode 0 --sched 0" \
            "--cols 64 --mode 0 --sched 1" "--cols 64 --mode 1" "--cols 64 --mode 0 --generic"; do
  echo "== $tool: kernel_bench $args" >> "$log"
  compute-sanitizer --tool "$tool" --error-exitcode 1 \
    python tools/kernel_bench.py --rows 60000 $args --iters 1 --warmup 1 --verify "${ext[@]}" >> "$log" 2>&1 || rc=1
done
echo "== $tool overall exit $rc" >> "$log"
exit $rc
