#!/bin/bash
# AddressSanitizer + UBSan, then ThreadSanitizer, over the host stages of `quant` (chunk index, the parallel parser with its 4-byte
# stores of 3-byte values, the text formatter) without a GPU: builds a small driver around afqh_host_stage_bench and runs it on
# reference-layout fixtures (plain / extra tags / 16-base UMIs; records of 1, 7, 9 and 300 alignments at chunk starts and ends).
# usage: scripts/host_sanitize.sh   (prints the sanitizer reports, if any; silent runs end with "clean")
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, os, numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(sys.argv[0])) if False else os.getcwd())
sys.path.insert(0, os.path.join(root, "tests")); sys.path.insert(0, root)
import rad_fixture
from test_host_cli import _fixture_cells
rng = np.random.default_rng(3)
n_refs = 5000; names = [f"t{i}" for i in range(n_refs)]
for tag, umi_len, kw in (("plain", 12, {}), ("extra", 12, dict(extra_read_tags=[(("frag_q", "u8"), 7)], extra_aln_tags=[(("pos", "u32"), 5)], extra_first=True)), ("wide", 16, {})):
    cells = _fixture_cells(rng, 300, n_refs, 16, umi_len, min_recs=1, max_recs=80)
    for c in (0, 3, 11, 299):
        bc, recs = cells[c]
        for na in (7, 300, 9, 1):
            refs = sorted(set(int(x) for x in rng.integers(0, n_refs, size=na)))
            recs.insert(len(recs) if na in (9, 1) else 1, (int(rng.integers(0, 1 << 20)), refs, [True] * len(refs)))
    rad_fixture.write_collated_rad(os.path.join(sys.argv[1], tag + ".rad"), names, cells, 16, umi_len, **kw)
PY
cat > $T/main.cpp <<EOF2
#include <cstdio>
#include "$ROOT/include/afq_host.h"
int main(int argc, char** argv) {
  int bad = 0;
  for (int i = 1; i < argc; ++i)
    for (int fe = 0; fe < 2; ++fe) {
      afqh_stage_info s{}; char err[512] = {0};
      if (afqh_host_stage_bench(argv[i], 4, fe ? 3 : 0, &s, err, sizeof err) != 0) { printf("%s: %s\n", argv[i], err); bad = 1; }
    }
  return bad;
}
EOF2
SRC="$T/main.cpp $ROOT/alevin_fry_b200/host/host_quant.cpp $ROOT/alevin_fry_b200/host/rad.cpp $ROOT/alevin_fry_b200/host/snappy_frame.cpp"
LNK="-L$ROOT/alevin_fry_b200 -lafq -lz -Wl,-rpath,$ROOT/alevin_fry_b200 -pthread"
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -o $T/asan $SRC $LNK
g++ -O1 -g -std=c++17 -fsanitize=thread -fno-omit-frame-pointer -o $T/tsan $SRC $LNK
for env in "" "AFQ_NO_PACK24=1"; do
  env $env ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 $T/asan $T/plain.rad $T/extra.rad $T/wide.rad
  env $env TSAN_OPTIONS=report_signal_unsafe=0 $T/tsan $T/plain.rad $T/extra.rad
done
rm -rf $T
echo clean
