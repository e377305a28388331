#!/bin/bash
# A/B of branch `elect-issue` (warp-uniform elect.sync issue blocks, NOTES.md section 4.0) against main in ONE gpurun call.
#
#   bash scripts/ab_elect.sh build        here, no GPU: builds the branch's library next to main's as
#                                         ttt_video_dit_b200/lib/libttt_b200_elect.so (git-ignored, travels with gpurun)
#   gpurun --timeout 900 -- 'bash scripts/ab_elect.sh run'
#                                         on the box: GPU parity suite + bench + secondary benches with each library
#                                         (TTT_B200_LIB selects it); results in gpurun_out/ab_{main,elect}_*.log
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
case "${1:-}" in
build)
  WT="$(mktemp -d)/wt"
  git worktree add -q "$WT" elect-issue || exit 1
  OBJ="$(mktemp -d)"
  objs=()
  for f in "$WT"/ttt_video_dit_b200/csrc/*.cu; do
    o="$OBJ/$(basename "${f%.cu}").o"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I"$WT/include" -c "$f" -o "$o" &
    objs+=("$o")
  done
  wait
  nvcc -shared -o ttt_video_dit_b200/lib/libttt_b200_elect.so "${objs[@]}" -gencode arch=compute_100a,code=sm_100a -lcuda 2>/dev/null \
    || nvcc -shared -o ttt_video_dit_b200/lib/libttt_b200_elect.so "${objs[@]}" -gencode arch=compute_100a,code=sm_100a
  git worktree remove --force "$WT"
  ls -la ttt_video_dit_b200/lib/
  ;;
run)
  mkdir -p gpurun_out
  for v in main elect; do
    lib="$ROOT/ttt_video_dit_b200/lib/libttt_b200.so"
    [ "$v" = elect ] && lib="$ROOT/ttt_video_dit_b200/lib/libttt_b200_elect.so"
    echo "=== $v ($lib)"
    TTT_B200_LIB="$lib" timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee "gpurun_out/ab_${v}_pytest.log"
    TTT_B200_LIB="$lib" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > "gpurun_out/ab_${v}_bench.json"
    cut -c1-200 "gpurun_out/ab_${v}_bench.json"
    TTT_B200_LIB="$lib" timeout 600 python scripts/bench_extra.py 2>&1 | grep -v Warn | tail -8 > "gpurun_out/ab_${v}_extra.log"
    cut -c1-220 "gpurun_out/ab_${v}_extra.log"
  done
  ;;
*)
  echo "usage: $0 build|run"; exit 2 ;;
esac
